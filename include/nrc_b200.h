/*
 * nrc_b200.h -- C ABI of the B200-native Neural-Radiance-Cache MLP (drop-in for VkNRC's hot path).
 *
 * The reference has no FFI boundary for this path: the MLP sits behind (i) the `VkNRCState` object
 * (src/VkNRCState.hpp:17-90) plus the descriptor-binding tables of its render-graph passes
 * (src/rg/NNInference.cpp:13-48, src/rg/NNTrain.cpp:59-66, 91-104) and (ii) the vuda `launchKernel` calls of the
 * test harness (test/main.cpp:116-117, 180-181). Each entry point below names the reference interface it replaces.
 *
 * Conventions: every function returns 0 on success and a negative NRC_ERR_* otherwise (never throws, never aborts);
 * nrc_last_error() returns a thread-local description of the last failure. All `d_*` pointers are CUDA device
 * pointers on the handle's device; `stream` is a cudaStream_t passed as void*. Work is enqueued, not synchronised,
 * and counts are read on the device (the reference uses indirect dispatch, src/rg/NNDispatch.hpp:22-44), so a frame
 * never needs a host round trip. A handle is not thread-safe (the reference records one command buffer per frame
 * on one thread, src/main.cpp:160-169).
 *
 * Buffer layouts are bit-identical to the reference's (see vknrc_b200/csrc/nrc_config.h):
 *   weights / use_weights  fp16 row-major W[l][out][in], layer l at l*4096, layer 5 = 3x64 at 20480 (NN_nv.glsl:69-82)
 *   optimizer_entries      {m, v, weight, ema_weight} fp32 (src/VkNRCState.cpp:29-31)
 *   optimizer_state        {u32 t; f32 beta1_t, beta2_t, alpha_t, alpha_t_1}   (src/VkNRCState.cpp:25-28)
 *   eval / train records   NRCEvalRecord 20 B, NRCTrainRecord 40 B             (shader/src/NRCRecord.glsl:6-38)
 *   gradients              fp32, weight layout, + [20672] loss sum, [20673] record count, padded to 20736
 */
#ifndef NRC_B200_H
#define NRC_B200_H
#include <stdint.h>

#include "nrc_b200_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NRC_OK 0
#define NRC_ERR_INVALID_ARGUMENT (-1)
#define NRC_ERR_CUDA (-2)
#define NRC_ERR_UNSUPPORTED_DEVICE (-3) /* not an sm_100 part: there is no fallback path */
#define NRC_ERR_OUT_OF_MEMORY (-4)
#define NRC_ERR_PEER_TIMEOUT (-5) /* multi-GPU: a peer's gradient words never arrived (nrc_comm_status) */

#define NRC_B200_WEIGHT_COUNT 20672u
#define NRC_B200_GRADIENT_FLOATS 20736u
#define NRC_B200_GRAD_LOSS_SLOT 20672u
#define NRC_B200_GRAD_COUNT_SLOT 20673u

typedef struct nrc_state_t *nrc_handle_t;

typedef struct nrc_config_t {
	uint32_t extent_width, extent_height; /* VkNRCState(queue, extent), src/VkNRCState.hpp:37 */
	uint64_t seed;                        /* explicit: the reference seeds from std::random_device */
} nrc_config_t;

const char *nrc_last_error(void);

/* ---- VkNRCState statics (src/VkNRCState.hpp:83-89, src/VkNRCState.cpp:34-37) ---- */
uint64_t nrc_get_eval_record_buffer_size(uint32_t extent_width, uint32_t extent_height);
uint64_t nrc_get_batch_train_record_buffer_size(void);
uint32_t nrc_get_train_batch_count(void);
uint32_t nrc_get_train_batch_size(void);
uint32_t nrc_get_weight_count(void);
float nrc_get_default_train_probability(void);

/* ---- VkNRCState object (src/VkNRCState.hpp:37-81) ---- */
int nrc_create(const nrc_config_t *config, int device, nrc_handle_t *out);                 /* ctor, :37-40 */
void nrc_destroy(nrc_handle_t h);
int nrc_reset_mlp_buffers(nrc_handle_t h, uint64_t seed);  /* ResetMLPBuffers, src/VkNRCState.cpp:46-88 (He-normal) */
int nrc_set_weights(nrc_handle_t h, const float *host_fp32_weights); /* same upload path, caller-provided values */
void *nrc_get_weight_buffer(nrc_handle_t h);               /* GetWeightBuffer,          fp16 x 20672 (device) */
void *nrc_get_use_weight_buffer(nrc_handle_t h);           /* GetUseWeightBuffer,       fp16 x 20672 (device) */
void *nrc_get_optimizer_entry_buffer(nrc_handle_t h);      /* GetOptimizerEntryBuffer,  16 B x 20672 (device) */
void *nrc_get_optimizer_state_buffer(nrc_handle_t h);      /* GetOptimizerStateBuffer,  20 B (device) */
void *nrc_get_gradient_buffer(nrc_handle_t h);             /* `gradients` of NNTrain (src/rg/NNTrain.hpp:96-98): fp32 x 20736 */
int nrc_download(nrc_handle_t h, uint16_t *weights, uint16_t *use_weights, void *optimizer_entries,
                 void *optimizer_state, float *gradients, void *stream); /* any pointer may be NULL; synchronises */
void nrc_set_use_ema_weights(nrc_handle_t h, int use_ema); /* SetUseEMAWeights */
int nrc_is_use_ema_weights(nrc_handle_t h);
void nrc_set_train_probability(nrc_handle_t h, float p);   /* SetTrainProbability (consumed by the record producer) */
float nrc_get_train_probability(nrc_handle_t h);
uint32_t nrc_next_frame(nrc_handle_t h);                   /* NextFrame: advances and returns the per-frame seed */
uint32_t nrc_get_seed(nrc_handle_t h);

/* ---- test-harness kernels on pre-encoded inputs ----
 * nrc_mlp_evaluate_encoded == launchKernel("evaluate_32.spv", ..., weights, inputs, outputs)   (test/main.cpp:116-117)
 * nrc_mlp_gradient_encoded == launchKernel("train_32.spv", ..., weights, dw, inputs, targets)  (test/main.cpp:180-181)
 * inputs [n][64] fp16, outputs/targets [n][3] fp16, dw fp32 x 20672 ACCUMULATED into (L2 loss, test/train_NV.comp).
 * n need not be a multiple of 128 (the reference's test kernels require it). */
int nrc_mlp_evaluate_encoded(const void *d_weights, const void *d_inputs, void *d_outputs, uint64_t n, void *stream);
int nrc_mlp_gradient_encoded(const void *d_weights, float *d_dw, const void *d_inputs, const void *d_targets,
                             uint64_t n, void *stream);

/* ---- inference (nrc_inference.comp; bindings src/rg/NNInference.cpp:13-48), weights = use_weights ----
 * `d_count` may be NULL (then max_count queries run). */
int nrc_infer_encoded(nrc_handle_t h, const void *d_inputs, void *d_outputs_f16vec3, uint64_t n, int clamp_output,
                      void *stream);
/* The same call on HOST buffers: what the reference's harness does around one launch - cudaMemcpy inputs up, launch,
 * cudaMemcpy outputs down (test/main.cpp:103-128) - as one call. The queries are cut in chunks whose host->device copy,
 * MLP and device->host copy overlap on three streams; the work is ordered after what `stream` holds at the call and
 * `stream` completes when the outputs are in `h_outputs`. Pinned host memory makes the copies truly asynchronous. */
int nrc_infer_encoded_host(nrc_handle_t h, const void *h_inputs, void *h_outputs_f16vec3, uint64_t n, int clamp_output,
                           void *stream);
/* The reference's own query format from HOST memory: `n` 20-byte NRCEvalRecords (the buffer path_tracer.comp fills, binding 8 of
 * nrc_inference.comp; `dst` is ignored) -> UnpackNRCInput + encode + MLP with use_weights -> max(y, 0) as fp16 x 3 per query in
 * `h_outputs`, same chunked three-stream pipeline: 20 bytes per query travel up and 6 down instead of 128 + 6. */
int nrc_infer_eval_records_host(nrc_handle_t h, const void *h_eval_records, uint64_t n, const NrcScene *scene,
                                void *h_outputs_f16vec3, void *stream);
/* records = 14 fp32 each (UnpackedNRCInput order, NRCRecord.glsl:40-45) `stride_bytes` apart; outputs max(y,0) fp16x3 */
int nrc_infer_unpacked(nrc_handle_t h, const void *d_records, uint32_t stride_bytes, const uint32_t *d_count,
                       uint64_t max_count, void *d_outputs_f16vec3, void *stream);
/* full output stage of nrc_inference.comp:48-73: dst words (NRCRecord.glsl:19-33) `dst_stride_bytes` apart select
 * screen (bias_factor_r rgba32f RMW + factor_gb rg32f, `image_pitch` pixels per row) or train-record feedback. */
int nrc_infer_scatter_unpacked(nrc_handle_t h, const uint32_t *d_dst, uint32_t dst_stride_bytes, const void *d_records,
                               uint32_t stride_bytes, const uint32_t *d_count, uint64_t max_count, void *d_bias_factor_r,
                               const void *d_factor_gb, uint32_t image_pitch, void *const d_train_records[4], void *stream);

/* ---- the reference's own record formats: NRCEvalRecord / NRCTrainRecord + scene buffers ----
 * nrc_infer == the nrc_inference.comp pass with the bindings of src/rg/NNInference.cpp:13-48: 20-byte eval records
 *   (binding 8) + device-resident count (9) + scene buffers / textures (0-7) -> UnpackNRCInput -> encode -> MLP with
 *   use_weights (10) -> max(y,0) -> scatter into bias_factor_r (11, rgba32f RMW) / factor_gb (12) or into the four
 *   batch_train_records buffers (13). `scene` is a HOST struct of device pointers.
 * nrc_infer_packed: same unpack + network on PackedNRCInput words `stride_bytes` apart, outputs max(y,0) as fp16x3.
 * nrc_train_batch == one NNTrain pass group (clear -> prepare -> gradient -> optimize) on a 40-byte NRCTrainRecord
 *   buffer (bindings src/rg/NNTrain.cpp:59-66, 91-104): target = record.bias, relative-L2-luminance loss, Adam + EMA.
 * nrc_train_frame == the frame's four pass groups (src/rg/NRCRenderGraph.cpp:57-70) in ONE kernel launch.
 * nrc_gradient: gradient + reduction only (for a caller that all-reduces the gradient buffer itself). */
int nrc_infer(nrc_handle_t h, const void *d_eval_records, const uint32_t *d_count, uint64_t max_count,
              const NrcScene *scene, void *d_bias_factor_r, const void *d_factor_gb, uint32_t image_pitch,
              void *const d_train_records[4], void *stream);
int nrc_infer_packed(nrc_handle_t h, const void *d_packed_inputs, uint32_t stride_bytes, const uint32_t *d_count,
                     uint64_t max_count, const NrcScene *scene, void *d_outputs_f16vec3, void *stream);
/* UnpackNRCInput on its own (NRCRecord.glsl:98-125): PackedNRCInput words -> [n][14] fp32 in UnpackedNRCInput order. */
/* Flattens the scene's index buffers into NrcScene::prim_table rows (64 B per primitive, 64-byte aligned, caller-owned
 * device memory of nrc_scene_prim_table_bytes(prim_count)). Once per scene: the reference never edits geometry after
 * load (src/VkScene.cpp:64-133). `scene->prim_table` itself is ignored here. */
uint64_t nrc_scene_prim_table_bytes(uint32_t prim_count);
int nrc_scene_build_prim_table(const NrcScene *scene, uint32_t prim_count, void *d_prim_table, void *stream);
int nrc_unpack_inputs(const void *d_packed_inputs, uint32_t stride_bytes, uint64_t n, const NrcScene *scene,
                      float *d_unpacked14, void *stream);
/* NRCInputEncode on its own (NRCRecord.glsl:77-95): [n] records of 14 fp32 in UnpackedNRCInput order (position, scattered_dir,
 * normal, roughness, diffuse, specular; `stride_bytes` >= 56 apart) -> [n][64] fp16 rows, the input layout of
 * test/evaluate_NV.comp / nrc_infer_encoded / nrc_gradient_encoded. Same device function as the fused paths: bit-identical
 * features. The *_packed form runs UnpackNRCInput (:98-125) in front of it, straight from PackedNRCInput words
 * (`stride_bytes` 16 for bare inputs, 20 / 40 inside NRCEvalRecord / NRCTrainRecord buffers offset to their packed_input). */
int nrc_encode_inputs(const void *d_inputs14, uint32_t stride_bytes, uint64_t n, void *d_encoded_f16x64, void *stream);
int nrc_encode_packed_inputs(const void *d_packed_inputs, uint32_t stride_bytes, uint64_t n, const NrcScene *scene,
                             void *d_encoded_f16x64, void *stream);
int nrc_gradient(nrc_handle_t h, const void *d_train_records, uint32_t *d_count, uint32_t max_count,
                 const NrcScene *scene, void *stream);
int nrc_train_batch(nrc_handle_t h, const void *d_train_records, uint32_t *d_count, uint32_t max_count,
                    const NrcScene *scene, int write_use_weights, void *stream);
int nrc_train_frame(nrc_handle_t h, void *const d_train_records[4], uint32_t *const d_counts[4], uint32_t max_count,
                    const NrcScene *scene, void *stream);

/* ---- one frame (src/rg/NRCRenderGraph.cpp) ----
 * nrc_frame_begin == NRCRenderGraph::PreExecute's counter reset (:108-112): eval_count and batch_train_count[0..3] <- 0,
 *   enqueued on `stream` BEFORE the caller's record producer (the path tracer) appends this frame's records.
 * nrc_frame == the NN part of the graph in its order (:46-80): nn_inference_pass with last frame's use_weights (screen
 *   composite + feedback into the train targets), then the four nn_train_pass groups on the counts the producer left,
 *   use_weights republished by the last one (if its batch is not empty, nrc_optimize.comp:33-34). Two kernel launches,
 *   no host synchronisation: the pair nrc_frame_begin / nrc_frame can be captured into a CUDA graph once and replayed. */
int nrc_frame_begin(nrc_handle_t h, uint32_t *d_eval_count, uint32_t *const d_train_counts[4], void *stream);
int nrc_frame(nrc_handle_t h, const void *d_eval_records, const uint32_t *d_eval_count, uint64_t max_eval_count,
              const NrcScene *scene, void *d_bias_factor_r, const void *d_factor_gb, uint32_t image_pitch,
              void *const d_train_records[4], uint32_t *const d_train_counts[4], void *stream);

/* ---- training (NNTrain pass group: clear -> prepare -> gradient -> optimize, src/rg/NNTrain.hpp:95-127) ----
 * nrc_gradient_unpacked: gradient pass + deterministic batch reduction into the gradient buffer (replaces clear +
 *   nrc_gradient.comp's atomics). inputs as above; targets = 3 fp32 (`bias`, NRCRecord.glsl:36) `target_stride` apart.
 *   `d_count` (optional) is clamped in place to max_count like nrc_train_prepare.comp:17-19.
 * nrc_adam_step: nrc_train_prepare.comp:22-28 + nrc_optimize.comp:32-54 using gradient[20673] as the record count
 *   (so that a multi-GPU caller can all-reduce the gradient buffer in between); writes `weights`, and `use_weights`
 *   iff write_use_weights (the reference does that for the frame's last batch only, NRCRenderGraph.cpp:66-68).
 * nrc_train_batch_unpacked = both in ONE kernel launch (the optimizer runs on the reduced gradient inside the same
 *   cooperative kernel). */
int nrc_gradient_unpacked(nrc_handle_t h, const void *d_inputs, uint32_t input_stride, const void *d_targets,
                          uint32_t target_stride, uint32_t *d_count, uint32_t max_count, void *stream);
int nrc_gradient_encoded(nrc_handle_t h, const void *d_inputs, const void *d_targets_f16vec3, uint32_t *d_count,
                         uint32_t max_count, int relative_loss, void *stream);
int nrc_adam_step(nrc_handle_t h, int write_use_weights, void *stream);
int nrc_train_batch_unpacked(nrc_handle_t h, const void *d_inputs, uint32_t input_stride, const void *d_targets,
                             uint32_t target_stride, uint32_t *d_count, uint32_t max_count, int write_use_weights,
                             void *stream);
/* nrc_train_frame_unpacked = the whole NNTrain schedule of one frame (src/rg/NRCRenderGraph.cpp:57-70): four dependent
 *   batches, each clear -> prepare -> gradient -> optimize, use_weights written by the last one, in ONE kernel launch. */
int nrc_train_frame_unpacked(nrc_handle_t h, const void *const d_inputs[4], uint32_t input_stride,
                             const void *const d_targets[4], uint32_t target_stride, uint32_t *const d_counts[4],
                             uint32_t max_count, void *stream);
/* optional fp32 [max_count][3] buffer that receives the (unclamped) training predictions of the next gradient call */
void nrc_set_prediction_capture(nrc_handle_t h, float *d_predictions);

/* ---- multi-GPU (records of every batch sharded per GPU) ----
 * The reference is single-GPU; this is the exchange step of the data-parallel path. After set-up every nrc_gradient_* /
 * nrc_train_* call all-reduces the reduced gradient (dW, loss sum, record count) with the peers INSIDE the training
 * kernel - {epoch, value} words pushed into every peer's inbox over NVLink, fixed rank-order sum, so the result is
 * bit-identical on all ranks - and the optimizer step that follows is replicated. All ranks must make the same sequence
 * of training calls (max_count may differ per rank). Two ways to set up the inboxes:
 *   (a) one process per GPU, CUDA IPC: nrc_comm_init allocates this rank's inbox and returns its 64-byte IPC handle; the
 *       caller gathers the handles of all ranks (rank order) through its own channel and passes them to nrc_comm_connect
 *       on every rank (then synchronises the ranks once). One unicast store per peer and word.
 *   (b) caller-owned buffers, nrc_comm_attach: `d_inboxes[r]` = rank r's buffer of nrc_comm_buffer_bytes() as addressable
 *       from this device - cudaDeviceEnablePeerAccess pointers in a single process that drives several GPUs, a
 *       cuMemCreate / NCCL-window / torch-symmetric-memory allocation across processes. `d_multicast` (optional) is the
 *       same set of buffers bound to one NVSwitch multicast object (cuMulticastCreate): the push is then ONE multimem.st
 *       per word, replicated by the switch (NVLS). The local buffer is zeroed by the call; synchronise the ranks after it.
 * A peer that does not show up within the timeout (nrc_comm_set_timeout, in polls of one word; default 2^24, about a
 * second) does not hang or kill the GPU context: the waiting rank skips that batch's optimizer step and raises a flag
 * that nrc_comm_status (synchronises `stream`) reports as NRC_ERR_PEER_TIMEOUT. */
uint32_t nrc_comm_handle_bytes(void);
int nrc_comm_init(nrc_handle_t h, uint32_t rank, uint32_t world, void *out_handle);
int nrc_comm_connect(nrc_handle_t h, const void *all_handles);
uint64_t nrc_comm_buffer_bytes(void);
int nrc_comm_attach(nrc_handle_t h, uint32_t rank, uint32_t world, void *const *d_inboxes, void *d_multicast);
int nrc_comm_status(nrc_handle_t h, void *stream);
void nrc_comm_set_timeout(nrc_handle_t h, uint32_t polls);
int nrc_comm_shutdown(nrc_handle_t h);
uint32_t nrc_comm_world(nrc_handle_t h);

/* ---- learn-an-image harness (test/mlp_learning_an_image/{gradient,optimize,inference}.comp) ----
 * one training step: `batch` random uv samples (pcg2d stream of push-constant seeds), one-blob-32 encoding,
 * bilinear RGBA8 target, L2 loss, SGD lr on the fp32 master weights; inference over a width x width grid to RGBA8. */
int nrc_image_train_step(nrc_handle_t h, const void *d_image_rgba8, uint32_t image_w, uint32_t image_h, uint32_t seed_x,
                         uint32_t seed_y, uint32_t batch, float lr, void *stream);
int nrc_image_infer(nrc_handle_t h, void *d_out_rgba8, uint32_t width, void *stream);

/* ---- Vulkan interop (optional) ----
 * The reference owns every buffer on this path as a VkBuffer / VkImage: the persistent MLP buffers in VkNRCState
 * (src/VkNRCState.cpp:59-88), the per-frame eval / train record, count and screen-image resources in the render graph
 * (src/rg/NRCRenderGraph.cpp:139-175). A renderer that keeps allocating them in Vulkan exports the VkDeviceMemory as an
 * opaque fd (VK_KHR_external_memory_fd, VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT) and imports it here; the mapped
 * device pointer is what the d_* arguments above take. Queue ordering (the render graph's barriers between
 * path_tracer.comp, the NN passes and screen.frag) becomes a timeline semaphore exported with
 * VK_KHR_external_semaphore_fd: wait(value) on the library's stream before nrc_infer / nrc_train_frame, signal(value+1)
 * after. A successfully imported fd belongs to the CUDA driver (do not close it). Mapped pointers are released with
 * cudaFree before nrc_external_memory_release. NOT EXECUTED in this repo's environment (no Vulkan loader / ICD): only the
 * argument and error paths are tested. */
typedef struct nrc_external_memory_t *nrc_external_memory_handle_t;
typedef struct nrc_external_semaphore_t *nrc_external_semaphore_handle_t;
int nrc_import_vulkan_memory_fd(int device, int fd, uint64_t allocation_size, int dedicated, nrc_external_memory_handle_t *out);
int nrc_external_memory_map_buffer(nrc_external_memory_handle_t m, uint64_t offset, uint64_t size, void **d_ptr);
int nrc_external_memory_release(nrc_external_memory_handle_t m);
int nrc_import_vulkan_timeline_semaphore_fd(int device, int fd, nrc_external_semaphore_handle_t *out);
int nrc_external_semaphore_wait(nrc_external_semaphore_handle_t s, uint64_t value, void *stream);
int nrc_external_semaphore_signal(nrc_external_semaphore_handle_t s, uint64_t value, void *stream);
int nrc_external_semaphore_release(nrc_external_semaphore_handle_t s);

#ifdef __cplusplus
}
#endif
#endif /* NRC_B200_H */
