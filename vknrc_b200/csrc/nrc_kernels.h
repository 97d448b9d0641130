// nrc_kernels.h -- parameter blocks and launchers shared by the CUDA kernels and the host object (internal).
#pragma once
#include <cuda.h> // CUtensorMap (type only; the encode entry point is resolved at run time, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>

#include "nrc_config.h"
#include "sm100_ptx.cuh"

#ifndef NRC_INFER_SLOTS
#define NRC_INFER_SLOTS 5 // 128-sample tiles in flight per SM: 5 x (64 accumulator + 32 operand) = 480 of 512 TMEM columns
#endif
// record input modes: slots + producer warps (unpack / encode ahead of the slots) share the 32 warps of a CTA.
// Measured at 1080p (14-float records): 5 slots + 12 producers 81.0 us, 4 + 16: 83.6, 4 + 12: 82.8, 4 + 14: 83.8; before the
// producer warps existed (5 slots doing their own unpack + encode, old encoder): 113 us.
#ifndef NRC_INFER_SLOTS_REC
#define NRC_INFER_SLOTS_REC 5
#endif
#ifndef NRC_INFER_PRODUCER_WARPS
#define NRC_INFER_PRODUCER_WARPS 12
#endif
#ifndef NRC_INFER_FREE_BACKOFF
#define NRC_INFER_FREE_BACKOFF 200 // ns between polls of a buffer's use counter
#endif
// packed records (scene gather): the producers are latency-bound, so they get more of the CTA's 32 warps
// (nrc_infer at 1080p: 3 slots + 20 producers 133 us, 3 + 16: 141, 4 + 16: 143, 2 + 24: 151)
#ifndef NRC_INFER_SLOTS_PACKED
#define NRC_INFER_SLOTS_PACKED 3
#endif
#ifndef NRC_INFER_PRODUCER_WARPS_PACKED
#define NRC_INFER_PRODUCER_WARPS_PACKED 20
#endif

namespace nrc {

enum InMode : int {
	NRC_IN_ENCODED = 0,   // [n][64] fp16, test/evaluate_NV.comp:18-21
	NRC_IN_UNPACKED = 1,  // [n] NrcUnpackedInput-shaped records (14 fp32 at a caller-given stride), encode fused
	NRC_IN_IMAGE_GRID = 2, // learn-an-image inference: uv from the pixel index (inference.comp:33-34)
	NRC_IN_IMAGE_RANDOM = 3, // learn-an-image training: uv from pcg2d(seed + gid) (gradient.comp:47-48)
	NRC_IN_PACKED = 4      // [n] PackedNRCInput (16 B inside 20 B eval / 40 B train records): UnpackNRCInput + encode fused
};
enum OutMode : int {
	NRC_OUT_F16VEC3 = 0, // [n][3] fp16, test/evaluate_NV.comp:29-30
	NRC_OUT_SCATTER = 1, // nrc_inference.comp:48-73
	NRC_OUT_RGBA8 = 2    // mlp_learning_an_image/inference.comp:53
};
enum LossKind : int { NRC_LOSS_L2 = 0, NRC_LOSS_RELATIVE_L2_LUMINANCE = 1 };

struct InferParams {
	uint64_t n;               // number of queries (upper bound when d_count != nullptr)
	const uint32_t *d_count;  // optional device-resident count
	int in_mode, out_mode, clamp_output;
	const void *in;           // NRC_IN_UNPACKED: first record's 14 floats; NRC_IN_PACKED: first record's PackedNRCInput
	uint32_t in_stride_bytes; // bytes between records (56 for an array of unpacked inputs, 20 for NRCEvalRecord)
	NrcScene scene;           // NRC_IN_PACKED: the buffers UnpackNRCInput gathers from
	uint32_t image_width;     // NRC_IN_IMAGE_GRID
	void *out;                // F16VEC3 / RGBA8
	const uint32_t *dst;      // SCATTER: eval-record dst words
	uint32_t dst_stride_u32;  // SCATTER: stride between dst words in u32 (5 inside an NrcEvalRecord array, 1 if packed)
	void *bias_factor_r;      // SCATTER: rgba32f image, read-modify-write
	const void *factor_gb;    // SCATTER: rg32f image
	uint32_t image_pitch;     // SCATTER: pixels per image row
	void *train_records[NRC_TRAIN_BATCH_COUNT];
};

struct GradParams {
	uint64_t n;              // records in this batch (upper bound when d_count != nullptr)
	uint32_t *d_count;       // optional device-resident count (read clamped to n; written back clamped to batch_cap)
	int in_mode, loss_kind;
	float loss_scale;
	const void *in;          // ENCODED: unused (TMA); UNPACKED: 14 floats per record at in_stride_bytes; PACKED: PackedNRCInput
	uint32_t in_stride_bytes;
	NrcScene scene;          // PACKED: the buffers UnpackNRCInput gathers from
	const void *target;      // ENCODED: [n][3] fp16 (test/train_NV.comp:11-16); UNPACKED: 3 fp32 at target_stride_bytes
	uint32_t target_stride_bytes;
	int target_is_f16;
	// learn-an-image (gradient.comp:46-50)
	uint32_t seed_x, seed_y;
	const uint8_t *image_rgba8;
	uint32_t image_w, image_h;
	float *partials;         // [gridDim.x][NRC_GRAD_STRIDE] per-CTA partial dW (+ loss, count slots)
	void *y_out;             // optional [n][3] fp32 predictions (unclamped), for loss-curve checks
};

struct AdamParams {
	const float *gradients;          // [NRC_GRAD_STRIDE]; the divisor is gradients[NRC_GRAD_COUNT_SLOT]
	NrcOptimizerEntry *entries;
	NrcOptimizerState *opt_state;    // advanced in place (nrc_train_prepare.comp:22-28) by the last CTA to finish
	uint32_t *done_counter;          // zero-initialised device word owned by the state object
	__half *weights;                 // fp16, reference layout
	__half *use_weights;             // nullptr = the non-WRITE_USE_WEIGHTS variant
	int use_ema;
};

// Multi-GPU exchange (one process per GPU, or one process driving several GPUs; records of every batch sharded per
// GPU): after the local reduction each rank pushes its reduced 64-float blocks into every peer's inbox over NVLink as
// 8-byte {epoch, value} words - ONE multimem.st through the NVSwitch multicast mapping of the inboxes (NVLS) when the
// caller attached one, else one st.relaxed.sys per peer into peer-mapped memory. One traversal carries data and flag,
// no fence, the receiver spins on the very word it needs and adds the R copies in rank order: an all-reduce fused into
// the training kernel, bit-identical on every rank, followed by a replicated Adam step. Each rank's record count
// travels the same way in ONE word per source (written by the rank's CTA 0, polled by every CTA of the receiver), so
// nothing in the layout depends on a rank's grid size.
// Comm buffer of a rank (identical layout everywhere, zero-initialised; parity = epoch & 1 double-buffers it):
//   uint64_t data [2 parities][NRC_MAX_RANKS sources][NRC_GRAD_STRIDE]        (epoch << 32 | float bits)
//   uint64_t count[2 parities][NRC_MAX_RANKS sources]                         (epoch << 32 | count float bits)
#define NRC_MAX_RANKS 8
struct CommParams {
	uint32_t rank, world;           // world <= 1: no exchange
	uint32_t *epoch_word;           // device word: epochs used so far; batch b of a launch uses epoch *epoch_word + 1 + b (never 0)
	uint64_t *inbox[NRC_MAX_RANKS]; // comm buffer of every rank as mapped into this process (inbox[rank] = the local one)
	uint64_t *multicast;            // the same buffers through one multicast mapping (stores land in every rank's inbox), or nullptr
	uint32_t *error_word;           // device word, set to 1 when a peer's word did not arrive within spin_limit polls
	uint32_t spin_limit;            // polls of one word before giving up (the launch then finishes without an optimizer step)
};
constexpr size_t kCommDataWords = 2ull * NRC_MAX_RANKS * NRC_GRAD_STRIDE;
constexpr size_t kCommCountWords = 2ull * NRC_MAX_RANKS;
constexpr size_t kCommBytes = (kCommDataWords + kCommCountWords) * sizeof(uint64_t);

// One launch of nrc_train_kernel = up to NRC_TRAIN_BATCH_COUNT dependent training batches (a frame):
// per batch: gradient pass (per-CTA partial dW) -> grid barrier -> deterministic reduction of the partials
// (+ optionally the optimizer step on the reduced gradient) -> grid barrier -> next batch with the new weights.
struct TrainParams {
	GradParams batch[NRC_TRAIN_BATCH_COUNT];
	uint32_t num_batches;
	int adam_mode[NRC_TRAIN_BATCH_COUNT]; // 0: reduce only, 1: + Adam/EMA step, 2: + also write use_weights
	float *gradients;      // [NRC_GRAD_STRIDE] reduced dW (+ loss, count slots) of the LAST reduced batch
	int accumulate;        // 1: gradients += sum (test/train_NV.comp semantics), 0: gradients = sum
	uint32_t limit;        // number of leading elements to produce (20672 for a caller's dW, NRC_GRAD_STRIDE otherwise)
	uint32_t batch_cap;    // d_count is clamped in place to this (nrc_train_prepare.comp:17-19)
	AdamParams adam;       // use_weights / use_ema are taken from here when adam_mode == 2
	uint32_t *grid_bar;    // {monotonic arrival counter of the grid barrier, its value before the next launch} (owned by the state)
	CommParams comm;
	uint32_t pool_tiles;   // activation tiles in shared memory (set by launch_train: 8, or 6 when no CTA has a second tile)
};

struct SgdParams { // mlp_learning_an_image/optimize.comp:21-29
	const float *gradients;
	NrcOptimizerEntry *entries; // fp32 master weights live in entries[i].weight
	__half *weights;
	float lr, batch;
};

cudaError_t launch_infer(const InferParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream);
// cooperative launch: grid = min(#tiles of the largest batch, #SMs) CTAs, all co-resident
cudaError_t launch_train(const TrainParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream, uint32_t *grid_out = nullptr);
cudaError_t launch_unpack(const void *packed, uint32_t stride_bytes, uint64_t n, const NrcScene &scene, float *out14, cudaStream_t stream);
// NRCInputEncode as a kernel of its own: in_mode NRC_IN_UNPACKED (14-float records) or NRC_IN_PACKED (PackedNRCInput + scene) -> [n][64] fp16
cudaError_t launch_encode(const void *in, int in_mode, uint32_t stride_bytes, uint64_t n, const NrcScene &scene, void *out, cudaStream_t stream);
cudaError_t launch_prim_table(const NrcScene &scene, uint32_t prim_count, void *rows, cudaStream_t stream);
cudaError_t launch_adam(const AdamParams &p, cudaStream_t stream);
cudaError_t launch_sgd(const SgdParams &p, cudaStream_t stream);
uint32_t gradient_max_partials(int sms);
constexpr uint32_t kMaxTrainGrid = 160; // the in-kernel reduction sums at most 16 groups x 10 partials

} // namespace nrc
