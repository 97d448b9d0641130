// nrc_train.cu -- online training of the NRC MLP on sm_100a: forward, loss gradient, back-propagation, batch
// weight-gradient reduction and the optimizer step, ONE cooperative kernel launch per batch or per frame. Replaces
//   shader/src/nrc_gradient.comp:26-58, test/train_NV.comp:18-46, test/mlp_learning_an_image/gradient.comp:46-80
//   (NN_nv.glsl: NNForward*, NNLoadDA3_*, NNBackwardDA*_ReLU, NNUpdateDW*),
//   shader/src/nrc_train_prepare.comp:16-28 + nrc_optimize.comp:32-54, mlp_learning_an_image/optimize.comp:21-29,
//   and the clear -> prepare -> gradient -> optimize pass group of src/rg/NNTrain.hpp:95-127 (x4 per frame,
//   src/rg/NRCRenderGraph.cpp:57-70).
//
// nrc_train_kernel, per CTA (one 128-record tile at a time; 8 epilogue warps = 4 TMEM lane quarters x 2 column
// halves, one issue warp whose elected thread issues every TMA load and tcgen05.mma, and 3 producer warps that unpack +
// encode the records of the tile after next into a spare activation buffer - a whole round ahead of the forward pass):
//   every operand tile is an array of 128-byte rows (64 fp16) in shared memory with the 128-byte swizzle:
//     W_l      [out][in]     used K-major  (forward B)   and MN-major (dA: B with K = out)
//     a_l      [sample][in]  used K-major  (forward A)   and MN-major (dW: B with K = sample)
//     delta_l  [sample][out] used K-major  (dA: A)       and MN-major (dW: A with K = sample)
//   so each tensor is stored once and read through two UMMA descriptor flavours - no transposes, no copies.
//   forward  l=0..5 : D[128 x 64|16] = a_l * W_l^T                  (M=128)  -> ReLU -> a_{l+1}: fp16 back into TENSOR memory as the
//                     next layer's A operand (TS-form MMAs; layer 0 reads a_0 from shared memory), plus a shared-memory copy for dW
//   loss            : delta_5 = dL/dy (NN_nv.glsl:162-196)
//   backward l=5..1 : D[128 x 64]   = delta_l * W_l                 (M=128, A from tensor memory)  -> * [a_l > 0] -> delta_{l-1}
//   dW       l=5..0 : dW_l[64 x 64] += delta_l^T * a_l  (K = 128 samples, M=64, both operands from shared memory) accumulated IN TMEM
//                     across all of the CTA's tiles (two M=64 accumulators share 64 columns at lane offsets 0 / 16: 192 columns),
//                     written once per CTA as a partial.
//   Hand-offs are mbarriers only: af_ready / ab_ready (8 warp arrivals: TMEM operand stored + accumulator drained -> issuer),
//   ds_ready (the shared-memory copy of delta_l is complete -> dW_l), df_full / db_full (tcgen05.commit -> epilogue warps),
//   tile_done (all MMAs of the CTA's last tile of a batch complete).
// Then, in the same launch: grid barrier -> the partials are summed in a fixed order (deterministic, unlike the
// reference's 2.6 M fp32 atomics per batch, NN_nv.glsl:309-314,357-364), each CTA owning a slice of the 20 736 floats,
// and (optionally) nrc_optimize.comp is applied verbatim to that slice -> grid barrier -> next batch of the frame.
#include "nrc_kernels.h"
#include <atomic>
#include "nrc_encode.cuh"
#include "nrc_unpack.cuh"
#include <type_traits>

using namespace sm100;

#ifdef NRC_TRACE
// development aid (tools/trace_grad.cu): thread 0 of CTA 0 logs (tag, clock) pairs straight into global memory (fire-and-forget
// stores; the kernel's shared memory is full)
#define NRC_GTRACE_CAP 512
__device__ uint2 g_nrc_gtrace[NRC_GTRACE_CAP];
__device__ unsigned int g_nrc_gtrace_n;
#define NRC_GTRACE(tag)                                                                                                \
	do {                                                                                                               \
		if (blockIdx.x == 0 && threadIdx.x == 0 && gtrace_n < NRC_GTRACE_CAP)                                          \
			g_nrc_gtrace[gtrace_n++] = make_uint2((uint32_t)(tag), (uint32_t)clock64());                               \
	} while (0)
// the issuing thread's own timeline (tags 0x100 + ...): merged with the epilogue thread's by time stamp when dumped
__device__ uint2 g_nrc_itrace[NRC_GTRACE_CAP];
__device__ unsigned int g_nrc_itrace_n;
#define NRC_ITRACE(tag)                                                                                                \
	do {                                                                                                               \
		if (blockIdx.x == 0 && itrace_n < NRC_GTRACE_CAP)                                                              \
			g_nrc_itrace[itrace_n++] = make_uint2((uint32_t)(tag), (uint32_t)clock64());                               \
	} while (0)
// one producer thread's timeline (tags 0x200 + ...)
__device__ uint2 g_nrc_ptrace[NRC_GTRACE_CAP];
__device__ unsigned int g_nrc_ptrace_n;
#define NRC_PTRACE(tag)                                                                                                \
	do {                                                                                                               \
		if (blockIdx.x == 0 && threadIdx.x == kProducerWarp0 * 32 && ptrace_n < NRC_GTRACE_CAP)                        \
			g_nrc_ptrace[ptrace_n++] = make_uint2((uint32_t)(tag), (uint32_t)clock64());                               \
	} while (0)
#elif defined(NRC_GTRACE_FENCE) // experiment: the trace points as pure compiler scheduling fences
#define NRC_GTRACE(tag) asm volatile("" ::: "memory")
#define NRC_ITRACE(tag)
#define NRC_PTRACE(tag)
#else
#define NRC_GTRACE(tag)
#define NRC_ITRACE(tag)
#define NRC_PTRACE(tag)
#endif

#ifndef NRC_COMM_POLL_PARALLEL
#define NRC_COMM_POLL_PARALLEL 0
#endif

namespace nrc {

namespace {
constexpr uint32_t kWOff = 0;                         // 6 x 8 KB weights
constexpr uint32_t kPoolOff = NRC_LAYERS * 8192;      // P x 16 KB activation tiles (also fp32 staging of the partial): P = 9 when a CTA
                                                      // runs several tiles (two in flight + the input tile of the one after), 6 when every
                                                      // CTA has at most one - the smaller footprint leaves the L1 more, which the
                                                      // latency-bound frame feels (-2 us)
// then 2 x 16 KB deltas (ping-pong: delta_l lives in buffer (5 - l) & 1; scratch of the reduction phase), then the barriers
constexpr uint32_t kPoolMulti = 9, kPoolSingle = 6;
constexpr uint32_t train_smem_bytes(uint32_t pool_tiles) { return kPoolOff + (pool_tiles + 2) * 16384 + 256 + 1024; }
static_assert(train_smem_bytes(kPoolMulti) + 256 <= 232448, "dynamic + static shared memory of a CTA");
// TMEM map (columns x lanes): the M=64 dW accumulators occupy 16 of every 32 lanes, so two of them share 64 columns - dW_l at
// columns 64*(l/2), lane offset 16*(l%2) (dW_4 with dW_5^T) - 192 columns instead of 336. That leaves room for the fp16 A
// operands of both streams in TMEM: forward (a_k) and back-propagation (delta_l) run TS-form, 32 instead of 50.8 cycles per
// MMA, and - more important - the next MMA of a stream only waits for a tcgen05.st, not for shared-memory stores + proxy fence.
constexpr uint32_t kColWorkF = 192, kColWorkB = 256;  // working accumulators of the forward / backward stream
constexpr uint32_t kColAF = 320, kColAB = 352;        // fp16 A operands (32 columns = 64 fp16 per lane): a_k / delta_l
__device__ __forceinline__ constexpr uint32_t dw_col(int l) { return l == 5 ? 128u : 64u * (uint32_t)(l >> 1); }
__device__ __forceinline__ constexpr uint32_t dw_lane(int l) { return l == 5 ? 16u : 16u * (uint32_t)(l & 1); }
constexpr uint32_t kEpiWarps = 8, kEpiThreads = 256, kIssueWarp = 8;
constexpr uint32_t kProducerWarp0 = 9, kProducerWarps = 3; // record input modes: raw record -> a_0, a tile ahead of the forward pass
constexpr int kMainThreads = 288;                           // warps 0..8: every CTA-wide barrier of the batch loop (named barrier 2)
constexpr int kTrainThreads = 384;
constexpr uint32_t kReduceBlocks = NRC_GRAD_STRIDE / 64; // the reduction works on blocks of 64 consecutive floats
static_assert(NRC_GRAD_STRIDE % 64 == 0, "the reduction works on 64-float blocks");
// Activation tiles of a CTA that runs several tiles per batch live in a ring of nine 16 KB buffers. Round r carries the forward
// pass of tile r (a_0..a_5 in fw[]) and the backward pass of tile r - 1 (bw[]): the forward tile holds a_0..a_{k+1} and the
// backward tile a_0..a_l at step k (l = 5 - k), (k + 2) + (l + 1) = 8 buffers, and the ninth (nx) receives a_0 of tile r + 1
// while the round runs. A buffer is handed over the moment its last reader is done:
//   a_{k+1} of a tile takes the buffer of a_{6-k} of the tile before it (k >= 1; freed by dW_{6-k} one step earlier),
//   a_1 takes the buffer of a_0 two tiles back (freed by dW_0 at the end of the previous round),
//   the next spare buffer is the one of a_1 two tiles back (freed by dW_1, step 4 of the previous round).
// The issuing thread, the epilogue warps and the producer warps each step their own copy; it restarts with every batch.
struct TileRing {
	uint32_t fw[6] = {0, 1, 2, 3, 4, 5}, bw[6] = {7, 6, 0, 0, 0, 0}, nx = 8;
	__device__ __forceinline__ void rotate() {
		const uint32_t n0 = nx, n1 = bw[0], n2 = fw[5], n3 = fw[4], n4 = fw[3], n5 = fw[2], spare = bw[1];
#pragma unroll
		for (int i = 0; i < 6; ++i)
			bw[i] = fw[i];
		fw[0] = n0, fw[1] = n1, fw[2] = n2, fw[3] = n3, fw[4] = n4, fw[5] = n5, nx = spare;
	}
};

// CTA-wide barriers of the training kernel's batch loop: warps 0..8 only (named barrier 2). The producer warps run ahead of
// the batch loop on their own schedule - they encode the next batch's first tile while these warps reduce the current
// batch's gradient - and meet the others again at the kernel's final __syncthreads.
__device__ __forceinline__ void main_sync() { asm volatile("bar.sync 2, %0;" ::"n"(kMainThreads) : "memory"); }
__device__ __forceinline__ bool main_sync_or(bool pred) {
	uint32_t r;
	asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbarrier.red.or.pred q, 2, %2, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
	             : "=r"(r)
	             : "r"((uint32_t)pred), "n"(kMainThreads)
	             : "memory");
	return r != 0;
}
} // namespace

// bilinear RGBA8 fetch, clamp-to-edge, normalised coordinates (the sampler of mlp_learning_an_image/main.cpp:121-124)
__device__ __forceinline__ void sample_bilinear_rgb(const uint8_t *img, uint32_t w, uint32_t h, float u, float v, float rgb[3]) {
	const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
	const float fx = floorf(x), fy = floorf(y);
	const float tx = x - fx, ty = y - fy;
	const int x0 = min(max((int)fx, 0), (int)w - 1), x1 = min(max((int)fx + 1, 0), (int)w - 1);
	const int y0 = min(max((int)fy, 0), (int)h - 1), y1 = min(max((int)fy + 1, 0), (int)h - 1);
	const uchar4 *p = (const uchar4 *)img;
	const uchar4 c00 = p[(size_t)y0 * w + x0], c10 = p[(size_t)y0 * w + x1], c01 = p[(size_t)y1 * w + x0], c11 = p[(size_t)y1 * w + x1];
	const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
	const float s = 1.0f / 255.0f;
	rgb[0] = (w00 * c00.x + w10 * c10.x + w01 * c01.x + w11 * c11.x) * s;
	rgb[1] = (w00 * c00.y + w10 * c10.y + w01 * c01.y + w11 * c11.y) * s;
	rgb[2] = (w00 * c00.z + w10 * c10.z + w01 * c01.z + w11 * c11.z) * s;
}

// ------------------------------------------------------------------------------------------------------------------
// Optimizer math: nrc_train_prepare.comp:22-28 (running products) + nrc_optimize.comp:32-54 (Adam + EMA).
// Explicit round-to-nearest intrinsics: no FMA contraction, so every step rounds where the GLSL source rounds.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ NrcOptimizerState advance_state(const NrcOptimizerState old) {
	NrcOptimizerState st;
	st.t = old.t + 1;
	st.beta1_t = old.beta1_t * NRC_ADAM_BETA1;
	st.beta2_t = old.beta2_t * NRC_ADAM_BETA2;
	st.alpha_t_1 = old.alpha_t;
	st.alpha_t = old.alpha_t * NRC_EMA_ALPHA;
	return st;
}
__device__ __forceinline__ void adam_update(const AdamParams &a, uint32_t i, NrcOptimizerEntry e, float grad_sum, float count,
                                            const NrcOptimizerState &st) {
	float g = __fdiv_rn(grad_sum, count);
	if (NRC_LOSS_SCALE != 1.0f) // (x / 1 == x exactly: the second IEEE division of nrc_optimize.comp:36 is skipped, not approximated)
		g = __fdiv_rn(g, NRC_LOSS_SCALE);
	if (isnan(g) || isinf(g))
		g = 0.0f;
	e.m = __fadd_rn(__fmul_rn(NRC_ADAM_BETA1, e.m), __fmul_rn(1.0f - NRC_ADAM_BETA1, g));
	e.v = __fadd_rn(__fmul_rn(NRC_ADAM_BETA2, e.v), __fmul_rn(1.0f - NRC_ADAM_BETA2, __fmul_rn(g, g)));
	const float hm = __fdiv_rn(e.m, 1.0f - st.beta1_t), hv = __fdiv_rn(e.v, 1.0f - st.beta2_t);
	e.weight = __fsub_rn(e.weight, __fdiv_rn(__fmul_rn(NRC_LEARNING_RATE, hm), __fadd_rn(__fsqrt_rn(hv), NRC_ADAM_EPSILON)));
	const float eta_t = 1.0f - st.alpha_t, eta_t_1 = 1.0f - st.alpha_t_1;
	e.ema_weight = __fadd_rn(__fmul_rn(__fdiv_rn(1.0f - NRC_EMA_ALPHA, eta_t), e.weight),
	                         __fmul_rn(__fmul_rn(NRC_EMA_ALPHA, eta_t_1), e.ema_weight)); // sic
	a.entries[i] = e;
	a.weights[i] = __float2half_rn(e.weight);
	if (a.use_weights)
		a.use_weights[i] = __float2half_rn(a.use_ema ? e.ema_weight : e.weight);
}
// The last CTA to finish publishes the advanced state: every CTA has derived it from the old one before arriving, so
// no CTA can observe a half-updated state and no extra launch is needed.
// (the caller has put a CTA-wide barrier in front: every thread's entries / weights are written)
__device__ __forceinline__ void publish_state_if_last(const AdamParams &a, const NrcOptimizerState &st) {
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(a.done_counter, 1u) == gridDim.x - 1) {
			*a.opt_state = st;
			*a.done_counter = 0;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Grid barrier. All CTAs of the (cooperative) launch are co-resident. ONE monotonically increasing arrival counter,
// never reset: barrier k of a launch is passed when the counter reaches base + (k + 1) * gridDim.x, where `base` is the
// counter's value before the launch. The base lives in device memory next to the counter (bar[1]): every CTA reads it
// when the kernel starts and CTA 0 stores the final value after the launch's last barrier - by then every CTA has read
// the old one - so nothing about the barrier is baked into the launch parameters and a captured CUDA graph replays
// correctly. Wrap-around is harmless, the comparison is on the signed difference. An arrival is one fire-and-forget red; the waiters poll the
// counter itself, so a barrier costs fence + one-way red + one load round trip (1800 cycles; the earlier generation-word
// scheme - load generation, atomic with return, last arriver stores, the others poll - measured 3050).
// Release: the CTA barrier orders every thread's global writes before thread 0's __threadfence + arrival. Acquire:
// readers after the barrier use L2 loads (ld.cg / TMA / volatile), so no trailing fence. kProxyFence: the writes before
// the barrier are read through the async proxy (TMA) after it.
// ------------------------------------------------------------------------------------------------------------------
#ifndef NRC_GRID_SYNC_BACKOFF
#define NRC_GRID_SYNC_BACKOFF 50
#endif
template <bool kProxyFence> __device__ __forceinline__ void grid_sync(uint32_t *counter, uint32_t &target) {
	if (kProxyFence)
		asm volatile("fence.proxy.async;" ::: "memory");
	main_sync();
	if (threadIdx.x == 0) {
		target += gridDim.x;
		__threadfence();
		asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
		uint32_t seen;
		for (;;) {
			asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
			if ((int32_t)(seen - target) >= 0)
				break;
			__nanosleep(NRC_GRID_SYNC_BACKOFF); // 148 pollers on one L2 line slow the arrivals down: 2440 -> 2320 cycles
		}
	}
	main_sync();
}

__device__ __forceinline__ float ld_cg(const float *p) { // L2 only: the partials were written by other SMs in this launch
	float v;
	asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
	return v;
}

// 8-byte {epoch, payload} words exchanged with peer GPUs: a single store is atomic, so the receiver can spin on the word
__device__ __forceinline__ void st_peer(uint64_t *p, uint64_t v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// the same word through the NVSwitch multicast mapping: one store, replicated by the switch into every rank's inbox
__device__ __forceinline__ void st_multicast(uint64_t *p, uint64_t v) { asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// Bounded wait for a peer's word. A peer that never shows up (crashed, or entered the launch `spin_limit` polls late) must
// neither hang the GPU nor kill the context: the waiter raises the handle's error word (reported by nrc_comm_status as
// NRC_ERR_PEER_TIMEOUT), stops waiting for the rest of the launch, and its CTA skips the optimizer step of this batch.
__device__ __forceinline__ uint32_t wait_peer(const uint64_t *p, uint32_t epoch, const CommParams &comm, bool &timed_out) {
	uint64_t got = 0;
	if (!timed_out) {
		for (uint32_t spins = 0;; ++spins) {
			asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(p) : "memory");
			if ((uint32_t)(got >> 32) == epoch)
				return (uint32_t)got;
			if (spins > comm.spin_limit)
				break;
			if (spins > 16)
				__nanosleep(32);
		}
		timed_out = true;
		atomicExch(comm.error_word, 1u);
	}
	return 0u;
}
// The all-reduce over NVLink, fused into the reduction phase: push {epoch, value} words into every peer's inbox (one multicast
// store through the switch, or one store per peer), spin on the peers' words in the local inbox, add in rank order (identical
// operands in identical order on every rank => bit-identical sums). Only the MULTI instantiations of the kernel contain it: its
// registers must not weigh on the single-GPU path (+6 us per frame when they did). Returns true if this CTA gave up waiting for a peer.
__device__ __forceinline__ bool exchange_with_peers(const CommParams &comm, uint32_t epoch, bool mine_blk, uint32_t my_i, float count, bool &have_peer_counts,
                                                 uint32_t *peer_counts, float &sum, float &total_count) {
	const uint32_t world = comm.world, me = comm.rank, parity = epoch & 1u;
	const size_t dslot = (size_t)parity * NRC_MAX_RANKS * NRC_GRAD_STRIDE, cslot = kCommDataWords + (size_t)parity * NRC_MAX_RANKS;
	const uint64_t tag = (uint64_t)epoch << 32;
	bool timed_out = false;
	if (mine_blk) {
		const size_t at = dslot + (size_t)me * NRC_GRAD_STRIDE + my_i;
		if (comm.multicast) {
			st_multicast(comm.multicast + at, tag | __float_as_uint(sum));
		} else {
			for (uint32_t r = 0; r < world; ++r)
				if (r != me)
					st_peer(comm.inbox[r] + at, tag | __float_as_uint(sum));
		}
	}
	if (!have_peer_counts && threadIdx.x >= 192 && threadIdx.x < 192 + NRC_MAX_RANKS) { // one thread per peer: the record counts
		const uint32_t r = threadIdx.x - 192;
		if (blockIdx.x == 0) { // this rank's count: one word per source, published once per batch by CTA 0
			if (comm.multicast) {
				if (r == me)
					st_multicast(comm.multicast + cslot + me, tag | __float_as_uint(count));
			} else if (r < world && r != me) {
				st_peer(comm.inbox[r] + cslot + me, tag | __float_as_uint(count));
			}
		}
		peer_counts[r] = r < world && r != me ? wait_peer(comm.inbox[me] + cslot + r, epoch, comm, timed_out) : 0u;
	}
	if (mine_blk) {
#if NRC_COMM_POLL_PARALLEL
		// all peers' words requested at once and polled together, then added in rank order. Measured at 8 GPUs: no faster than
		// waiting for the peers one after the other (88.5 us per frame either way - the words of a 128-byte line arrive together,
		// so after the first wait the others hit), and 16 more live registers: off by default.
		const uint64_t *src = comm.inbox[me] + dslot + my_i;
		uint32_t val[NRC_MAX_RANKS];
		uint32_t pending = ((1u << world) - 1u) & ~(1u << me);
		for (uint32_t spins = 0; pending && !timed_out; ++spins) {
			uint64_t got[NRC_MAX_RANKS];
#pragma unroll
			for (uint32_t r = 0; r < NRC_MAX_RANKS; ++r)
				if (pending >> r & 1u)
					asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got[r]) : "l"(src + (size_t)r * NRC_GRAD_STRIDE) : "memory");
#pragma unroll
			for (uint32_t r = 0; r < NRC_MAX_RANKS; ++r)
				if ((pending >> r & 1u) && (uint32_t)(got[r] >> 32) == epoch)
					val[r] = (uint32_t)got[r], pending &= ~(1u << r);
			if (pending && spins > comm.spin_limit) {
				timed_out = true;
				atomicExch(comm.error_word, 1u);
			}
			if (pending && spins > 16)
				__nanosleep(32);
		}
		float tot = 0.0f;
#pragma unroll
		for (uint32_t r = 0; r < NRC_MAX_RANKS; ++r)
			if (r < world)
				tot += r == me ? sum : (pending >> r & 1u) ? 0.0f : __uint_as_float(val[r]);
		sum = tot;
#else
		float tot = 0.0f;
		for (uint32_t r = 0; r < world; ++r)
			tot += r == me ? sum : __uint_as_float(wait_peer(comm.inbox[me] + dslot + (size_t)r * NRC_GRAD_STRIDE + my_i, epoch, comm, timed_out));
		sum = tot;
#endif
	}
	const bool cta_timed_out = main_sync_or(timed_out);
	have_peer_counts = true;
	total_count = 0.0f;
	for (uint32_t r = 0; r < world; ++r)
		total_count += r == me ? count : __uint_as_float(peer_counts[r]); // integers < 2^24: exact
	return cta_timed_out;
}

__device__ __forceinline__ float4 ld_cg4(const float *p) {
	float4 v;
	asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
	return v;
}

// (168 registers: the register file is per SM sub-partition, 16384 each, and one of the four hosts 3 of the 9 warps)
template <int IN_MODE, bool MULTI>
__global__ void __launch_bounds__(kTrainThreads, 1)
    nrc_train_kernel(const __grid_constant__ TrainParams tp, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ float scratch[16];
	__shared__ uint32_t peer_counts[NRC_MAX_RANKS];
	// every batch's record count (nrc_train_prepare.comp:17-18: min(count, capacity)), read once in the prologue: nothing in this
	// launch changes a batch's count before that batch's own reduction phase (which only writes the clamped value back), and a
	// dependent global load at the top of every batch iteration stood in front of the weight re-staging
	__shared__ unsigned long long batch_n[NRC_TRAIN_BATCH_COUNT];
	uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *w_sm = smem + kWOff, *pool_sm = smem + kPoolOff, *delta_sm = pool_sm + tp.pool_tiles * 16384;
	// reduction scratch [block of the round][partial group][16 x float4 = 64 floats] = 12 KB: lives in the delta buffers, which
	// only the gradient phase uses (the producer warps, which may be a batch ahead, never touch them)
	float4(*red_sm)[16][16] = (float4(*)[16][16])delta_sm;
	uint64_t *bars = (uint64_t *)(delta_sm + 2 * 16384);
	uint64_t *w_full = bars, *in_full = bars + 1 /* [2]: tile t uses in_full[t & 1] */, *df_full = bars + 3, *db_full = bars + 4, *dw1_done = bars + 5;
	uint64_t *tile_done = bars + 6, *af_ready = bars + 7, *ab_ready = bars + 8, *d5_ready = bars + 9, *w_ready = bars + 10, *ds_ready = bars + 11;
	uint32_t *tmem_slot = (uint32_t *)(bars + 12);
	// hand-back of input buffers to the producer warps: (batch << 16 | rounds of the batch whose output-layer MMAs have completed),
	// written by epilogue thread 0 - a counter, not a parity: a waiter can never miss a phase
	volatile uint32_t *in_free = tmem_slot + 1;
#ifdef NRC_TRACE
	uint32_t gtrace_n = 0, itrace_n = 0, ptrace_n = 0;
#endif
	NRC_GTRACE(1);

	// epilogue warps 0..7: q = TMEM lane quarter (rows 32q..32q+31), h = which 32-column half of the 64-wide row
	// (warp index through a shuffle: provably warp-uniform, so TMEM addresses / descriptors live in uniform registers)
	const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31, q = warp & 3, h = (warp >> 2) & 1;
	const uint32_t row = q * 32 + lane;

	if (threadIdx.x == 0) {
		mbar_init(w_full, 1), mbar_init(df_full, 1), mbar_init(db_full, 1), mbar_init(dw1_done, 1), mbar_init(tile_done, 1);
		// an input tile arrives by TMA (one arrival + bytes) or as four quarter tiles from the producer warps
		mbar_init(in_full, IN_MODE == NRC_IN_ENCODED ? 1 : 4), mbar_init(in_full + 1, IN_MODE == NRC_IN_ENCODED ? 1 : 4);
		mbar_init(af_ready, kEpiWarps), mbar_init(ab_ready, kEpiWarps), mbar_init(d5_ready, kEpiWarps), mbar_init(w_ready, kEpiWarps);
		mbar_init(ds_ready, kEpiWarps);
		*in_free = 0u;
		fence_mbar_init();
	}
	if (threadIdx.x >= 32 && threadIdx.x < 32 + NRC_TRAIN_BATCH_COUNT) {
		const uint32_t bi = threadIdx.x - 32;
		unsigned long long cnt = 0;
		if (bi < tp.num_batches) {
			cnt = tp.batch[bi].n;
			if (tp.batch[bi].d_count) {
				const unsigned long long c = *(volatile uint32_t *)tp.batch[bi].d_count;
				cnt = c < cnt ? c : cnt;
			}
		}
		batch_n[bi] = cnt;
	}
	if (warp == kIssueWarp)
		tmem_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
	NRC_GTRACE(2);

	constexpr uint32_t id_fwd64 = make_idesc_f16_f32(128, 64, false, false);
	constexpr uint32_t id_fwd16 = make_idesc_f16_f32(128, 16, false, false);
	constexpr uint32_t id_da = make_idesc_f16_f32(128, 64, false, true);  // A = delta K-major, B = W MN-major
	constexpr uint32_t id_dw64 = make_idesc_f16_f32(64, 64, true, true);  // A = delta MN-major, B = act MN-major
	constexpr uint32_t id_dw5t = make_idesc_f16_f32(64, 16, true, true);  // dW_5^T: A = a_5 MN-major, B = delta_5 MN-major
	// UMMA descriptors differ only in the start-address field: desc(addr + off) = desc(addr) + (off >> 4)
	// (only the start-address field = the low word varies; see mma_ss_lh)
	const uint32_t w_desc = smem_desc_lo(smem_u32(w_sm)), pool_desc = smem_desc_lo(smem_u32(pool_sm)), del_desc = smem_desc_lo(smem_u32(delta_sm));
	constexpr uint32_t dhi = kSmemDescHiSw128;
	const uint32_t df_issue = tmem + kColWorkF, db_issue = tmem + kColWorkB; // issuer's view of the two working accumulators
	const uint32_t df_mine = tmem_addr(tmem, q * 32, kColWorkF + 32 * h), db_mine = tmem_addr(tmem, q * 32, kColWorkB + 32 * h);
	const uint32_t af_issue = tmem + kColAF, ab_issue = tmem + kColAB;           // issuer's view of the two TMEM A operands
	const uint32_t af_mine = tmem_addr(tmem, q * 32, kColAF + 16 * h), ab_mine = tmem_addr(tmem, q * 32, kColAB + 16 * h); // this thread's half row
	// operand stored in TENSOR memory + accumulator drained -> one arrival per warp (no shared-memory traffic on this path)
	auto arrive_tmem = [&](uint64_t *bar) {
		tc_wait_st();
		tc_fence_before();
		__syncwarp();
		if (lane == 0)
			mbar_arrive(bar);
	};

	// operand stored (generic-proxy smem writes fenced to the async proxy) + accumulator drained -> one arrival per warp
	auto arrive_ready = [&](uint64_t *bar) {
		fence_proxy_async_smem();
		tc_wait_st();
		tc_fence_before();
		__syncwarp();
		if (lane == 0)
			mbar_arrive(bar);
	};
	auto store_half_row = [&](uint8_t *tile, const uint32_t *o16) { // 16 packed pairs = 32 columns = 4 swizzled 16 B chunks
		uint8_t *r = tile + row * 128;
#pragma unroll
		for (int c = 0; c < 4; ++c)
			*(uint4 *)(r + (((4 * h + c) ^ (row & 7)) << 4)) = make_uint4(o16[4 * c], o16[4 * c + 1], o16[4 * c + 2], o16[4 * c + 3]);
	};

	// Producer warps: this lane's whole row of a_0 for record `gi` of batch `bp` (n = the batch's clamped record count) as 32
	// packed fp16 pairs. nrc_gradient.comp:27-34: an invalid record is a zero input with a zero target => exactly zero contribution
	auto encode_record = [&](const GradParams &bp, uint64_t n, uint64_t gi, uint32_t o[32]) {
#pragma unroll
		for (int i = 0; i < 32; ++i)
			o[i] = 0u;
		if (gi >= n)
			return;
		if (IN_MODE == NRC_IN_PACKED) { // nrc_gradient.comp:29-31: UnpackNRCInput, then the same encoding
			float in[14];
			uint32_t pk[4];
			load_packed_input(bp.in, gi, bp.in_stride_bytes, pk);
			unpack_nrc_input(bp.scene, pk, in);
			encode_nrc(in, o);
		} else if (IN_MODE == NRC_IN_UNPACKED) {
			float in[14];
			const float2 *src = (const float2 *)((const uint8_t *)bp.in + gi * bp.in_stride_bytes);
#pragma unroll
			for (int i = 0; i < 7; ++i) {
				const float2 t = __ldg(src + i);
				in[2 * i] = t.x, in[2 * i + 1] = t.y;
			}
			encode_nrc(in, o);
		} else if (IN_MODE == NRC_IN_IMAGE_RANDOM) { // gradient.comp:47-49
			uint32_t px = bp.seed_x + (uint32_t)(gi % 128u), py = bp.seed_y + (uint32_t)(gi / 128u);
			pcg2d(px, py);
			const float sc = 1.0f / (float)0xffffffffu;
			encode_oneblob32(sc * (float)px, sc * (float)py, o);
		}
	};
	auto batch_count = [&](uint32_t bi) -> uint64_t { return batch_n[bi]; };
	auto tiles_of_this_cta = [&](uint64_t n) -> uint32_t {
		const uint32_t ntiles = (uint32_t)((n + NRC_TILE - 1) / NRC_TILE);
		return blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	};

	// mbarrier phase counters (every barrier completes once per use; the waiter tracks the parity of its next wait).
	// Hand-offs: in_full[t & 1] = a_0 of the CTA's t-th tile of the batch is in shared memory (TMA, or the four quarter tiles of
	// the producer warps) -> forward MMAs 0; af_ready = a_k stored by forward epilogue k-1 -> forward MMAs k; ab_ready = delta_l
	// stored by backward epilogue l+1 -> backward MMAs l; d5_ready = delta_5 stored by the forward epilogue of the output
	// layer -> backward MMAs 5 of the next round (a barrier of its own: nothing orders it against the previous tile's
	// last ab_ready phase); df_full / db_full = accumulator ready; dw1_done (pre-encoded inputs) = dW_1 has finished
	// reading a_1, whose buffer the TMA load of the tile after next overwrites.
	uint32_t df_ph = 0, db_ph = 0;            // epilogue threads
	uint32_t af_ph = 0, ab_ph = 0, d5_ph = 0; // issuer
	uint32_t ds_ph = 0;                       // issuer (TS form): delta_l is also in shared memory (dW_l's operand)
	uint32_t in_ph = 0, dw1_ph = 0;           // issuer: input tiles (bit j: parity of in_full[j]), dw1_done
	uint32_t done_ph = 0;                     // epilogue threads: tile_done completes once per batch in which the CTA had tiles

	if (warp >= kProducerWarp0) {
		// ================================================================================== producer warps
		// Record input modes: unit u = quarter (u & 3) of the CTA's tile t = u >> 2 of a batch (32 records, one per lane): raw
		// record -> (UnpackNRCInput ->) the 64 encoded features -> the row's eight 16-byte chunks at their 128-byte-swizzle
		// positions, exactly what TMA writes for pre-encoded inputs. The warps take the units round-robin and run AHEAD of the
		// batch loop: tile t + 1 is produced during round t (forward of tile t, backward of tile t - 1) into the ring's spare
		// buffer, the first tile of batch b + 1 while the other warps reduce batch b's gradient - the gather / encode latency
		// (2 000 .. 6 000 cycles per tile) never sits between two rounds or in front of a grid barrier. They take no part in the
		// CTA-wide barriers of the batch loop. A buffer is handed back through the counter in_free (see above).
		if (IN_MODE != NRC_IN_ENCODED) {
			const uint32_t pw = warp - kProducerWarp0;
#pragma unroll 1
			for (uint32_t b = 0; b < tp.num_batches; ++b) {
				const GradParams &bp = tp.batch[b];
				const uint64_t nb = batch_count(b);
				const uint32_t tiles_b = tiles_of_this_cta(nb);
				TileRing ring;
#pragma unroll 1
				for (uint32_t t = 0; t < tiles_b; ++t) {
					const uint32_t buf = t == 0 ? ring.fw[0] : ring.nx;
					// tiles 0 and 1 go to buffers that are free once the previous batch's gradient phase (incl. the staged dW) is
					// over; tile t >= 2 takes the buffer dW_1 of round t - 2 has read last - and its in_full barrier is the one of
					// tile t - 2, which must have completed its phase before the first arrival of this one
					const uint32_t need = (b << 16) | (t < 2 ? 0u : t - 1u);
#pragma unroll 1
					for (uint32_t qtr = 0; qtr < 4; ++qtr) {
						if ((4 * t + qtr) % kProducerWarps != pw)
							continue;
						NRC_PTRACE(0x200 + t);
						const uint32_t prow = qtr * 32 + lane;
						uint32_t o[32];
						encode_record(bp, nb, (uint64_t)(blockIdx.x + t * gridDim.x) * NRC_TILE + prow, o);
						NRC_PTRACE(0x210 + t);
						for (uint32_t spins = 0; (int32_t)(ld_acquire_cta(in_free) - need) < 0; ++spins) {
							if (spins > (1u << 25))
								__trap(); // a protocol bug must not hang the GPU (several seconds: longer than any peer time-out)
							__nanosleep(100);
						}
						__threadfence_block();
						uint8_t *dst = pool_sm + buf * 16384 + prow * 128;
#pragma unroll
						for (int c = 0; c < 8; ++c)
							*(uint4 *)(dst + ((c ^ (prow & 7)) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
						fence_proxy_async_smem(); // generic-proxy stores -> visible to the MMA's async-proxy operand reads
						__syncwarp();
						if (lane == 0)
							mbar_arrive(in_full + (t & 1));
						NRC_PTRACE(0x220 + t);
					}
					if (t >= 1)
						ring.rotate();
				}
			}
		}
	} else {
	// ====================================================================================== warps 0..8: the batch loop
	uint64_t n = batch_count(0);
	uint32_t my_tiles = tiles_of_this_cta(n);
	if (my_tiles == 0 && warp < kEpiWarps) {
		// A CTA without a tile in batch 0 never issues the TMA weight load whose out-of-bounds fill pads W_5 from 3 to 64 rows
		// (rows 323..383 of the weight tile): it writes those zeros itself, once. The same threads stage the weights in the
		// first batch where the CTA has work, and their proxy fence there also covers these stores.
		for (uint32_t idx = NRC_WEIGHT_ROWS * 8 + threadIdx.x; idx < NRC_LAYERS * 64 * 8; idx += kEpiThreads)
			*(uint4 *)(w_sm + (idx >> 3) * 128 + (((idx & 7u) ^ ((idx >> 3) & 7u)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
	}
	uint32_t w_reloads = 0; // weight re-stagings by the epilogue warps so far (w_ready phase)
	uint32_t bar_target = *(volatile const uint32_t *)(tp.grid_bar + 1); // (meaningful in thread 0 only)
	// multi-GPU: the epoch of this launch's first exchange, also device-resident (advanced by CTA 0 below)
	const uint32_t epoch_base = MULTI && tp.comm.world > 1 ? *(volatile const uint32_t *)tp.comm.epoch_word + 1u : 0u;

#pragma unroll 1
	for (uint32_t b = 0; b < tp.num_batches; ++b) {
		const GradParams &p = tp.batch[b];
		float *my_partial = p.partials + (size_t)blockIdx.x * NRC_GRAD_STRIDE;
		// the next batch's record count is read now (nothing writes it before that batch's own reduction phase)
		uint64_t n_next = 0;
		uint32_t tiles_next = 0;
		if (b + 1 < tp.num_batches) {
			n_next = batch_count(b + 1);
			tiles_next = tiles_of_this_cta(n_next);
		}
		if (b > 0 && my_tiles && warp < kEpiWarps) {
			// Re-stage the weights the previous batch's optimizer phase just wrote (by other SMs; the grid barrier made them
			// visible in L2): plain 16-byte L2 loads into the swizzled tile, no global cross-proxy fence needed. Rows past 323
			// hold zeros: from TMA's out-of-bounds fill in batch 0, or written in the prologue by a CTA without a tile in batch 0.
			const uint4 *src = (const uint4 *)tp.adam.weights;
			constexpr int kPerThread = (NRC_WEIGHT_ROWS * 8 + kEpiThreads - 1) / kEpiThreads; // 11: every load in flight at once
			uint4 v[kPerThread];
#pragma unroll
			for (int u = 0; u < kPerThread; ++u) {
				const uint32_t idx = threadIdx.x + u * kEpiThreads;
				if (idx < NRC_WEIGHT_ROWS * 8)
					asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + idx));
			}
#pragma unroll
			for (int u = 0; u < kPerThread; ++u) {
				const uint32_t idx = threadIdx.x + u * kEpiThreads, r = idx >> 3, c = idx & 7u;
				if (idx < NRC_WEIGHT_ROWS * 8)
					*(uint4 *)(w_sm + r * 128 + ((c ^ (r & 7u)) << 4)) = v[u];
			}
			fence_proxy_async_smem();
			__syncwarp();
			if (lane == 0)
				mbar_arrive(w_ready);
			NRC_GTRACE(0x5A);
		}

		// ---------------------------------------------------------------------------------------------------------------
		// Gradient phase. The CTA's tiles run as a two-stage software pipeline over "rounds": round r carries the FORWARD
		// pass of tile r and the BACKWARD pass of tile r - 1, step k = 0..5 of a round pairing forward layer k with backward
		// layer l = 5 - k. The two streams use different working accumulators, so while the epilogue warps convert one
		// stream's accumulator the tensor pipe runs the other stream's MMAs.
		//   forward  step k : D_F = a_k W_k^T               -> a_{k+1} = relu(D_F)   (k = 5: prediction -> loss gradient delta_5)
		//   backward step l : D_B = delta_l W_l             -> delta_{l-1} = D_B * [a_l > 0]   (two delta buffers, ping-pong)
		//                     dW_l += delta_l^T a_l          (accumulates in TMEM across all tiles of the CTA)
		// Activation tiles: see TileRing. Every reuse of a buffer is ordered by a tcgen05.commit that was issued after the last
		// MMA reading the old contents.
		// ---------------------------------------------------------------------------------------------------------------
		TileRing ring;

		if (my_tiles == 0) {
			// nothing to do in the gradient phase: the reduction below only reads the partials of the CTAs that had a tile
		} else if (warp == kIssueWarp) {
			// ================================================================================== issue warp
			if (elect_one()) {
				if (b == 0) {
					tma_prefetch_desc(&tm_w);
					mbar_arrive_expect_tx(w_full, NRC_LAYERS * 8192);
					for (int l = 0; l < NRC_LAYERS; ++l)
						tma_load_2d(w_sm + l * 8192, &tm_w, 0, l * 64, w_full);
				}
				// pre-encoded inputs: a_0 of tile t arrives by TMA on in_full[t & 1] - tiles 0 and 1 right away, tile r + 1 at the
				// start of round r into the ring's spare buffer: a whole round ahead (a TMA load from HBM takes 1 500+ cycles; issued
				// at the end of the round before, as in the first version, the forward stream stood still for that long every round)
				auto load_input_tile = [&](uint32_t t, uint32_t buf) {
					mbar_arrive_expect_tx(in_full + (t & 1), 16384);
					tma_load_2d(pool_sm + buf * 16384, &tm_in, 0, (int32_t)((blockIdx.x + t * gridDim.x) * NRC_TILE), in_full + (t & 1));
				};
				if (IN_MODE == NRC_IN_ENCODED) {
					load_input_tile(0, ring.fw[0]);
					if (my_tiles > 1)
						load_input_tile(1, ring.nx);
				}
				if (b == 0)
					mbar_wait(w_full, 0);
				else
					mbar_wait(w_ready, w_reloads & 1);
				// (three instantiations - forward only / both / backward only - so that the one-tile-per-CTA case of the
				// paper-sized batch runs straight-line code instead of hopping over the other stream's half of every step)
				auto issue_round = [&](auto HF, auto HB, uint32_t r) {
					constexpr bool has_f = decltype(HF)::value, has_b = decltype(HB)::value;
#pragma unroll
					for (int k = 0; k < NRC_LAYERS; ++k) {
						const int l = 5 - k;
						if (has_b && k == 0) { // delta_5 stored - and the forward accumulator of the output layer drained by every warp
							mbar_wait(d5_ready, d5_ph);
							d5_ph ^= 1;
						}
						if (has_f) { // ---- forward layer k of tile r
							if (k == 0) { // a_0 of tile r: in shared memory, by TMA or from the producer warps
								mbar_wait(in_full + (r & 1), (in_ph >> (r & 1)) & 1u);
								in_ph ^= 1u << (r & 1);
							} else {
								mbar_wait(af_ready, af_ph);
								af_ph ^= 1;
							}
							tc_fence_after();
							NRC_ITRACE(0x160 + k);
							const uint32_t a_d = pool_desc + ring.fw[k] * (16384 >> 4), b_d = w_desc + (uint32_t)(k * (8192 >> 4));
							if (k > 0) { // a_k from tensor memory (a_0 is in shared memory: SS form)
#pragma unroll
								for (int kk = 0; kk < 4; ++kk)
									mma_ts_lh(df_issue, af_issue + kk * 8, b_d + kk * 2, dhi, k < 5 ? id_fwd64 : id_fwd16, kk > 0);
							} else
							{
#pragma unroll
								for (int kk = 0; kk < 4; ++kk)
									mma_ss_lh(df_issue, a_d + kk * 2, b_d + kk * 2, dhi, k < 5 ? id_fwd64 : id_fwd16, kk > 0);
							}
							tc_commit(df_full);
							NRC_ITRACE(0x170 + k);
							if (IN_MODE == NRC_IN_ENCODED && k == 0 && r >= 1 && r + 1 < my_tiles) {
								// a_0 of tile r + 1 -> the spare buffer: the one of a_1 of tile r - 2, which dW_1 of the previous round read last
								if (r >= 2) {
									mbar_wait(dw1_done, dw1_ph);
									dw1_ph ^= 1;
								}
								load_input_tile(r + 1, ring.nx);
							}
						}
						if (has_b) { // ---- backward layer l of tile r - 1: dA first (critical path), then dW_l
							if (l < 5) {
								mbar_wait(ab_ready, ab_ph); // delta_l stored
								ab_ph ^= 1;
							}
							tc_fence_after();
							NRC_ITRACE(0x180 + l);
							const uint32_t dl = del_desc + (uint32_t)(((5 - l) & 1) * (16384 >> 4));
							const uint32_t al = pool_desc + ring.bw[l] * (16384 >> 4), wl = w_desc + (uint32_t)(l * (8192 >> 4));
							const uint32_t acc = (r > 1) ? 1u : 0u; // dW accumulates from the CTA's second tile on
							const uint32_t dw_acc = tmem_addr(tmem, dw_lane(l), dw_col(l)); // dW_l's accumulator
							if (l == 5) {
								mma_ts_lh(db_issue, ab_issue, wl, dhi, id_da, 0);
								tc_commit(db_full);
								mbar_wait(ds_ready, ds_ph); // delta_5 has reached shared memory too
								ds_ph ^= 1;
								tc_fence_after();
#pragma unroll
								for (int kk = 0; kk < 8; ++kk)
									mma_ss_lh(dw_acc, al + kk * 128, dl + kk * 128, dhi, id_dw5t, acc | (kk > 0));
							} else {
								if (l > 0) {
#pragma unroll
									for (int kk = 0; kk < 4; ++kk)
										mma_ts_lh(db_issue, ab_issue + kk * 8, wl + kk * 128, dhi, id_da, kk > 0);
									tc_commit(db_full);
									NRC_ITRACE(0x190 + l);
									mbar_wait(ds_ready, ds_ph); // delta_l has reached shared memory too
									ds_ph ^= 1;
									tc_fence_after();
									NRC_ITRACE(0x1A0 + l);
								}
#pragma unroll
								for (int kk = 0; kk < 8; ++kk)
									mma_ss_lh(dw_acc, dl + kk * 128, al + kk * 128, dhi, id_dw64, acc | (kk > 0));
								NRC_ITRACE(0x1B0 + l);
								if (l == 0 && !has_f) // the CTA's last tile of this batch: every dW accumulator is final
									tc_commit(tile_done);
								// (only where the wait below follows: every completed phase of dw1_done is consumed, so the
								// waiter's parity can never drift from the barrier's - also across the batches of a launch)
								// (a_1's buffer becomes the spare of the next round, which the load of tile r + 2 fills)
								if (IN_MODE == NRC_IN_ENCODED && l == 1 && r + 2 < my_tiles)
									tc_commit(dw1_done);
							}
						}
					}
					ring.rotate();
				};
				issue_round(std::true_type{}, std::false_type{}, 0u);
#pragma unroll 1
				for (uint32_t r = 1; r < my_tiles; ++r)
					issue_round(std::true_type{}, std::true_type{}, r);
				issue_round(std::false_type{}, std::true_type{}, my_tiles);
			}
			__syncwarp();
		} else {
			// ================================================================================== epilogue warps
			float loss_acc = 0.0f;
			uint32_t valid_rows = 0;
			// Draining a finished dW accumulator (M=64 TMEM layout: row r -> lane (r%16) + 32*(r/16)): every warp on its own - its 16
			// rows x 32 columns go TMEM -> a 2 KB slice of a dead pool buffer (16-byte chunks XOR-swizzled per row against bank
			// conflicts) -> this CTA's partial, each store instruction covering four 128-byte row segments. No CTA-wide barrier: the
			// version that staged the whole layer and copied it with all 256 threads cost the backward chain of the CTA's last tile
			// ~650 cycles per layer - every batch of a paper-sized frame - because a warp could not return to the chain before the
			// slowest one had staged. (Unstaged stores straight from the registers were tried: 16-byte ones +4 us per frame - the
			// partial sectors make the release fence of the grid barrier slow -, 32-byte ones +0.8 us.) For all but the last two layers this runs inside the backward
			// pass of the CTA's last tile, in the time the epilogue warps would spend waiting for the next accumulator.
			auto drain_dw = [&](int l, float *stage) {
				uint32_t v[32];
				tmem_ld_x32(tmem_addr(tmem, q * 32, dw_col(l) + 32 * h), v);
				tc_wait_ld();
				float4 *mine = (float4 *)stage + (q * 2 + h) * 128;
				if ((lane & 16u) == dw_lane(l)) { // the 16 lanes of this quarter that hold dW_l's rows 16q .. 16q+15
					const uint32_t rr = lane & 15u;
#pragma unroll
					for (int i = 0; i < 8; ++i)
						mine[rr * 8 + (i ^ (rr & 7u))] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
						                                            __uint_as_float(v[4 * i + 3]));
				}
				__syncwarp();
				float4 *out4 = (float4 *)(my_partial + l * 4096 + q * 16 * 64 + 32 * h);
#pragma unroll
				for (int it = 0; it < 4; ++it) {
					const uint32_t idx = it * 32 + lane, rr = idx >> 3, c = idx & 7u;
					out4[rr * 16 + c] = mine[rr * 8 + (c ^ (rr & 7u))];
				}
			};
			// Two accumulators that share their TMEM columns (dW_le in lanes 0..15, dW_le+1 in lanes 16..31 of every quarter, le even)
			// drained by ONE tcgen05.ld: every lane stages a row, the warp copies both layers' slices.
			auto drain_two = [&](int le, float *stage_e, float *stage_o) {
				uint32_t v[32];
				tmem_ld_x32(tmem_addr(tmem, q * 32, dw_col(le) + 32 * h), v);
				tc_wait_ld();
				const uint32_t slice = (q * 2 + h) * 128, rr0 = lane & 15u;
				float4 *mine = (float4 *)((lane & 16u) ? stage_o : stage_e) + slice;
#pragma unroll
				for (int i = 0; i < 8; ++i)
					mine[rr0 * 8 + (i ^ (rr0 & 7u))] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
					                                              __uint_as_float(v[4 * i + 3]));
				__syncwarp();
#pragma unroll
				for (int o = 0; o < 2; ++o) {
					const float4 *src = (const float4 *)(o ? stage_o : stage_e) + slice;
					float4 *out4 = (float4 *)(my_partial + (le + o) * 4096 + q * 16 * 64 + 32 * h);
#pragma unroll
					for (int it = 0; it < 4; ++it) {
						const uint32_t idx = it * 32 + lane, rr = idx >> 3, c = idx & 7u;
						out4[rr * 16 + c] = src[rr * 8 + (c ^ (rr & 7u))];
					}
				}
			};
			uint32_t fin1 = 1, fin0 = 0;        // staging buffers of dW_1 / dW_0 (the last tile's a_1 / a_0 buffers)
			float tgt[3] = {0.0f, 0.0f, 0.0f}; // the forward tile's target (learn-an-image: sampled at step 0, used at step 5)
			// record modes: the target's raw bits - requested at step 0, converted where they are used (step 5). A conversion next to
			// the load made every epilogue warp sit out the load's whole latency at the start of each round (~1 000 cycles of 7 400)
			uint32_t traw[3] = {0u, 0u, 0u};
			bool valid_f = false;
			uint64_t gi_f = 0;
			auto epilogue_round = [&](auto HF, auto HB, uint32_t r) {
				constexpr bool has_f = decltype(HF)::value, has_b = decltype(HB)::value, last_round = !has_f;
				if (has_f) {
					const uint32_t tile = blockIdx.x + r * gridDim.x;
					gi_f = (uint64_t)tile * NRC_TILE + row;
					valid_f = gi_f < n;
					tgt[0] = tgt[1] = tgt[2] = 0.0f;
					traw[0] = traw[1] = traw[2] = 0u; // (+0.0 in either format)
					if (h == 0 && valid_f && IN_MODE != NRC_IN_IMAGE_RANDOM) { // loaded first: the latency hides under the forward pass
						if (p.target_is_f16) {
							const uint16_t *t = (const uint16_t *)((const uint8_t *)p.target + gi_f * p.target_stride_bytes);
							traw[0] = __ldg(t), traw[1] = __ldg(t + 1), traw[2] = __ldg(t + 2);
						} else {
							const uint32_t *t = (const uint32_t *)((const uint8_t *)p.target + gi_f * p.target_stride_bytes);
							traw[0] = __ldg(t), traw[1] = __ldg(t + 1), traw[2] = __ldg(t + 2);
						}
					}
					if (IN_MODE == NRC_IN_IMAGE_RANDOM && h == 0 && valid_f) { // target = the image at this sample's uv (gradient.comp:47-50)
						uint32_t px = p.seed_x + (uint32_t)(gi_f % 128u), py = p.seed_y + (uint32_t)(gi_f / 128u);
						pcg2d(px, py);
						const float sc = 1.0f / (float)0xffffffffu;
						sample_bilinear_rgb(p.image_rgba8, p.image_w, p.image_h, sc * (float)px, sc * (float)py, tgt);
					}
				}
				NRC_GTRACE(3);
#pragma unroll
				for (int k = 0; k < NRC_LAYERS; ++k) {
					const int l = 5 - k;
					if (has_f) { // ------------------------------------------------------------ forward epilogue, layer k
						mbar_wait(df_full, df_ph);
						df_ph ^= 1;
						tc_fence_after();
						NRC_GTRACE(0x10 + k);
						// the output layer's commit covers every MMA issued before it - dW_1 of this round's backward tile among
						// them: its a_1 buffer is the next spare one (and the in_full barrier of tile r has long completed its phase)
						if (IN_MODE != NRC_IN_ENCODED && k == 5 && threadIdx.x == 0)
							st_release_cta(in_free, (b << 16) | (r + 1u));
						if (k < NRC_HIDDEN_LAYERS) { // a_{k+1} = fp16(relu(D))
							uint32_t v[32], o[16];
							tmem_ld_x32(df_mine, v);
							tc_wait_ld();
#pragma unroll
							for (int i = 0; i < 16; ++i)
								o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
							tmem_st_x16(af_mine, o); // next layer's A operand: the forward chain only waits for this
							// the shared-memory copy (dW_{k+1}'s operand and the ReLU mask of the backward pass) is ISSUED under the latency of
							// that TMEM store (the arrival below waits for the TMEM store only, not for these stores; their first reader is
							// issued after later arrivals of this warp, which order them and their proxy fence): -2 .. 3 % on the 2^20 step
							store_half_row(pool_sm + ring.fw[k + 1] * 16384, o);
							arrive_tmem(af_ready);
							NRC_GTRACE(0x20 + k);
							// (with a backward epilogue following in this step, its arrive_ready fences both copies: a fence of its own here
							// made this warp sit out the drain of these stores before it could poll the backward accumulator's barrier)
							if (!(has_b && l >= 1))
								fence_proxy_async_smem();
							NRC_GTRACE(0x60 + k);
						} else { // output layer + loss gradient (NN_nv.glsl:148-196) -> delta_5
							if (h == 0) {
								uint32_t yv[4];
								tmem_ld_x4(tmem_addr(tmem, q * 32, kColWorkF), yv);
								tc_wait_ld();
								float y[3], g[3]; // y is the fp16 network output widened to fp32, as NNOutput3 returns it
#pragma unroll
								for (int c = 0; c < 3; ++c)
									y[c] = __half2float(__float2half_rn(__uint_as_float(yv[c])));
								float inv_den = 1.0f;
								if (p.loss_kind == NRC_LOSS_RELATIVE_L2_LUMINANCE) {
									const float lum = 0.299f * fmaxf(y[0], 0.0f) + 0.587f * fmaxf(y[1], 0.0f) + 0.114f * fmaxf(y[2], 0.0f);
									inv_den = __frcp_rn(lum * lum + 0.01f); // one correctly rounded reciprocal instead of six divisions
								}
#pragma unroll
								for (int c = 0; c < 3; ++c) {
									const float t = IN_MODE == NRC_IN_IMAGE_RANDOM ? tgt[c]
									                : p.target_is_f16      ? __half2float(__ushort_as_half((unsigned short)traw[c]))
									                                       : __uint_as_float(traw[c]);
									const float d = y[c] - t;
									g[c] = 2.0f * p.loss_scale * d * inv_den;
									if (valid_f)
										loss_acc += d * d * inv_den;
								}
								if (!valid_f)
									g[0] = g[1] = g[2] = 0.0f;
								valid_rows += valid_f ? 1u : 0u;
								if (valid_f && p.y_out) {
									float *yo = (float *)p.y_out + 3 * gi_f;
									yo[0] = y[0], yo[1] = y[1], yo[2] = y[2];
								}
								uint8_t *rr = delta_sm + row * 128; // delta_5 (buffer 0): 16 fp16 = logical chunks 0 and 1 of the row
								const uint32_t d5[8] = {cvt_pack_f16x2(g[0], g[1]), cvt_pack_f16x2(g[2], 0.0f), 0u, 0u, 0u, 0u, 0u, 0u};
								tmem_st_x8(tmem_addr(tmem, q * 32, kColAB), d5); // K = 16: the backward stream's first A operand
								*(uint4 *)(rr + ((0 ^ (row & 7)) << 4)) = make_uint4(cvt_pack_f16x2(g[0], g[1]), cvt_pack_f16x2(g[2], 0.0f), 0u, 0u);
								*(uint4 *)(rr + ((1 ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
							}
							NRC_GTRACE(0x25);
							arrive_ready(d5_ready); // delta_5 feeds backward layer 5 (step 0 of the next round)
							arrive_ready(ds_ready); // (both copies are complete here: the loss epilogue is not on a per-layer chain)
						}
					}
					if (has_b && l >= 1) { // --------------------------------------------------- backward epilogue, layer l
						// delta_{l-1} = fp16(D) * [a_l > 0], NaN -> 0 (NN_nv.glsl:198-220, 240-242)
						mbar_wait(db_full, db_ph);
						db_ph ^= 1;
						tc_fence_after();
						NRC_GTRACE(0x30 + l);
						uint32_t v[32], a[16], o[16];
						tmem_ld_x32(db_mine, v);
						{
							const uint8_t *rr = pool_sm + ring.bw[l] * 16384 + row * 128;
#pragma unroll
							for (int c = 0; c < 4; ++c) {
								const uint4 t = *(const uint4 *)(rr + (((4 * h + c) ^ (row & 7)) << 4));
								a[4 * c] = t.x, a[4 * c + 1] = t.y, a[4 * c + 2] = t.z, a[4 * c + 3] = t.w;
							}
						}
						tc_wait_ld();
#pragma unroll
						for (int i = 0; i < 16; ++i) {
							const uint32_t d2 = cvt_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
							const __half2 dh = *(const __half2 *)&d2, ah = *(const __half2 *)&a[i];
							o[i] = d2 & __hgt2_mask(ah, __float2half2_rn(0.0f)) & __heq2_mask(dh, dh);
						}
						if (l >= 2) { // delta_{l-1} feeds dA_{l-1} from tensor memory; the shared-memory copy (dW_{l-1}) follows off the chain
							tmem_st_x16(ab_mine, o);
							store_half_row(delta_sm + ((6 - l) & 1) * 16384, o); // (issued under the TMEM store's latency, see the forward epilogue)
							arrive_tmem(ab_ready);
							NRC_GTRACE(0x40 + l);
							fence_proxy_async_smem(); // the copy (and the forward copy of this step) -> async proxy, then the second arrival
							__syncwarp();
							if (lane == 0)
								mbar_arrive(ds_ready);
							NRC_GTRACE(0x70 + l);
						} else { // delta_0 only feeds dW_0
							store_half_row(delta_sm + ((6 - l) & 1) * 16384, o);
							NRC_GTRACE(0x40 + l);
							arrive_ready(ab_ready);
						}
						if (last_round && l <= 4) { // dW_{l+1} is final (its MMAs precede this step's commits): drain it now
							if (l == 4) {
								if (h == 0) {
									uint32_t v5[4];
									tmem_ld_x4(tmem_addr(tmem, q * 32, dw_col(5)), v5); // dW_5^T: lane <-> in, column <-> out
									tc_wait_ld();
									if ((lane & 16u) == dw_lane(5)) {
										float *dst = my_partial + 5 * 4096 + q * 16 + (lane & 15u);
										dst[0] = __uint_as_float(v5[0]), dst[64] = __uint_as_float(v5[1]), dst[128] = __uint_as_float(v5[2]);
									}
								}
							} else if (l == 3) { // dW_4 (its column partner dW_5^T went out a step ago); staging: the dead tile of a_4
								NRC_GTRACE(0x80 + l);
								drain_dw(4, (float *)(pool_sm + ring.bw[4] * 16384));
								NRC_GTRACE(0x90 + l);
							} else if (l == 1) { // dW_2 (final now) together with dW_3 (final since the step before): they share columns
								NRC_GTRACE(0x80 + l);
								drain_two(2, (float *)(pool_sm + ring.bw[2] * 16384), (float *)(pool_sm + ring.bw[3] * 16384));
								NRC_GTRACE(0x90 + l);
							}
						}
					}
				}
				if (last_round)
					fin1 = ring.bw[1], fin0 = ring.bw[0];
				ring.rotate();
			};
			epilogue_round(std::true_type{}, std::false_type{}, 0u);
#pragma unroll 1
			for (uint32_t r = 1; r < my_tiles; ++r)
				epilogue_round(std::true_type{}, std::true_type{}, r);
			epilogue_round(std::false_type{}, std::true_type{}, my_tiles);
			// ---- the last two layers' dW complete with tile_done (dW_5..dW_2 were drained during the backward pass); their
			// staging slices are in the last tile's a_1 / a_0 buffers (dead once every MMA has completed)
			NRC_GTRACE(5);
			mbar_wait(tile_done, done_ph);
			done_ph ^= 1;
			tc_fence_after();
			NRC_GTRACE(6);
			drain_two(0, (float *)(pool_sm + fin0 * 16384), (float *)(pool_sm + fin1 * 16384));
			// loss / count slots: fixed-order reduction over the four h == 0 warps (deterministic)
			float cnt = (float)valid_rows;
#pragma unroll
			for (int off = 16; off > 0; off >>= 1) {
				loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
				cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
			}
			if (h == 0 && lane == 0)
				scratch[q] = loss_acc, scratch[4 + q] = cnt;
			tc_fence_before();
			asm volatile("bar.sync 1, 256;" ::: "memory");
			NRC_GTRACE(7);
			if (threadIdx.x < (NRC_GRAD_STRIDE - NRC_WEIGHT_COUNT) / 4) { // loss, count, zero padding
				float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				if (threadIdx.x == 0)
					v.x = (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]), v.y = (scratch[4] + scratch[5]) + (scratch[6] + scratch[7]);
				((float4 *)my_partial)[NRC_WEIGHT_COUNT / 4 + threadIdx.x] = v;
			}
			NRC_GTRACE(8);
		}
		if (IN_MODE != NRC_IN_ENCODED && warp < kEpiWarps) {
			// this batch's gradient phase is over: the producer warps may fill the first input buffers of the next batch
			// (thread 0 is past the barrier above, which every epilogue warp reaches after its last drain: nothing reads the
			// pool buffers any more)
			if (threadIdx.x == 0)
				st_release_cta(in_free, (b + 1u) << 16);
		}
		w_reloads += (b > 0 && my_tiles) ? 1u : 0u;

		// ============================================================ deterministic reduction (+ optimizer step)
		NrcOptimizerState pending_state{};
		bool publish_pending = false;
		grid_sync<false>(tp.grid_bar, bar_target);
		NRC_GTRACE(9);
		if (blockIdx.x == 0 && threadIdx.x == 0) { // (every CTA has passed this launch's first barrier: the old values are read)
			if (b + 1 == tp.num_batches)
				tp.grid_bar[1] = bar_target; // the launch's last barrier: publish the counter's final value as the next base
			if (MULTI && b == 0 && tp.comm.world > 1)
				*tp.comm.epoch_word = epoch_base - 1u + tp.num_batches;
		}
		{
			// (every CTA with a tile in this batch wrote one partial: CTAs 0 .. min(grid, #tiles) - 1)
			const uint64_t batch_tiles = (n + NRC_TILE - 1) / NRC_TILE;
			const uint32_t num_partials = batch_tiles < gridDim.x ? (uint32_t)batch_tiles : gridDim.x;
			// the batch's record count: integers < 2^24, so any summation order is exact (the load is issued here, its
			// reduction happens below, under the latency of the partial loads)
			float cnt_part = threadIdx.x < num_partials ? ld_cg(p.partials + (size_t)threadIdx.x * NRC_GRAD_STRIDE + NRC_GRAD_COUNT_SLOT) : 0.0f;
			float count = 0.0f;
			bool have_count = false;
			NRC_GTRACE(0x50);
			const int adam_mode = tp.adam_mode[b];
			AdamParams adam = tp.adam;
			if (adam_mode != 2)
				adam.use_weights = nullptr;
			NrcOptimizerState st{};
			if (adam_mode != 0) {
				const volatile NrcOptimizerState *os = adam.opt_state; // rewritten by the previous batch of this launch
				NrcOptimizerState old;
				old.t = os->t, old.beta1_t = os->beta1_t, old.beta2_t = os->beta2_t, old.alpha_t = os->alpha_t, old.alpha_t_1 = os->alpha_t_1;
				st = advance_state(old);
			}
			const uint32_t world = MULTI ? tp.comm.world : 1u, epoch = epoch_base + b;
			// The 20 736 floats are cut in blocks of 64; CTA c owns blocks c, c + grid, ... and handles up to three of them
			// per round. Per block: 16 threads x float4 cover the 64 floats of one partial, 16 groups of them take the
			// partials g, g + 16, g + 32, ... (all loads in flight at once), and the 16 group sums are combined by a
			// fixed binary tree. Every order is fixed => bit-reproducible.
			const uint32_t lane16 = threadIdx.x & 15u, grp = (threadIdx.x >> 4) & 15u;
			bool any_adam = false, have_peer_counts = false;
			for (uint32_t blk0 = blockIdx.x; blk0 < kReduceBlocks; blk0 += 3 * gridDim.x) {
				// the optimizer entry of "my" element of this round is independent of the sums: fetch it first
				const uint32_t my_blk = blk0 + (threadIdx.x >> 6) * gridDim.x, my_i = my_blk * 64 + (threadIdx.x & 63u);
				const bool mine_blk = threadIdx.x < 192 && my_blk < kReduceBlocks, mine = mine_blk && my_i < tp.limit;
				float sum = 0.0f;
				NrcOptimizerEntry my_entry{};
				if (mine && adam_mode != 0 && my_i < NRC_WEIGHT_COUNT) { // (L2 read: rewritten by another SM earlier in this launch)
					const float4 ev = ld_cg4((const float *)(adam.entries + my_i));
					my_entry.m = ev.x, my_entry.v = ev.y, my_entry.weight = ev.z, my_entry.ema_weight = ev.w;
				}
				if (threadIdx.x < 256) {
					float4 v[3][10]; // every load of the round is issued before the first add
#pragma unroll
					for (uint32_t c = 0; c < 3; ++c) {
						const uint32_t blk = blk0 + c * gridDim.x;
						const float *src = p.partials + blk * 64 + lane16 * 4;
#pragma unroll
						for (int u = 0; u < 10; ++u) {
							const uint32_t pp = grp + 16 * u;
							v[c][u] = blk < kReduceBlocks && pp < num_partials ? ld_cg4(src + (size_t)pp * NRC_GRAD_STRIDE) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
						}
					}
					if (!have_count) { // block-wide sum of the per-CTA record counts (all threads < 256 take this path together)
#pragma unroll
						for (int off = 16; off > 0; off >>= 1)
							cnt_part += __shfl_xor_sync(0xffffffffu, cnt_part, off);
						if (lane == 0)
							scratch[warp] = cnt_part;
					}
#pragma unroll
					for (uint32_t c = 0; c < 3; ++c) {
						float4 acc = v[c][0];
#pragma unroll
						for (int u = 1; u < 10; ++u)
							acc.x += v[c][u].x, acc.y += v[c][u].y, acc.z += v[c][u].z, acc.w += v[c][u].w;
						red_sm[c][grp][lane16] = acc;
					}
				}
				NRC_GTRACE(0x51);
				main_sync();
				if (!have_count) {
					for (int w = 0; w < 8; ++w)
						count += scratch[w];
					have_count = true;
				}
				NRC_GTRACE(0x58);
				if (mine_blk) {
					const float *col = &red_sm[threadIdx.x >> 6][0][(threadIdx.x & 63u) >> 2].x + (threadIdx.x & 3u);
					float t[16];
#pragma unroll
					for (int k = 0; k < 16; ++k)
						t[k] = col[k * 16 * 4];
#pragma unroll
					for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
						for (int k = 0; k < w; ++k)
							t[k] = t[k] + t[k + w];
					NRC_GTRACE(0x52);
					sum = t[0];
				}
				float total_count = count;
				bool cta_timed_out = false;
				if (MULTI && world > 1)
					cta_timed_out = exchange_with_peers(tp.comm, epoch, mine_blk, my_i, count, have_peer_counts, peer_counts, sum, total_count);
				if (mine) {
					const bool do_adam = adam_mode != 0 && total_count > 0.0f && !cta_timed_out; // nrc_optimize.comp:33-34 / nrc_train_prepare.comp:22
					any_adam = any_adam || do_adam;
					tp.gradients[my_i] = tp.accumulate ? tp.gradients[my_i] + sum : sum;
					NRC_GTRACE(0x55);
					if (do_adam && my_i < NRC_WEIGHT_COUNT)
						adam_update(adam, my_i, my_entry, sum, total_count, st);
					NRC_GTRACE(0x56);
				}
				main_sync();
				NRC_GTRACE(0x57);
			}
			NRC_GTRACE(0x53);
			if (blockIdx.x == 0 && threadIdx.x == 0 && p.d_count) { // nrc_train_prepare.comp:17-19: write the clamped count back
				const uint32_t cc = *p.d_count;
				*p.d_count = cc < tp.batch_cap ? cc : tp.batch_cap;
			}
			if (adam_mode != 0 && main_sync_or(any_adam)) { // (the batch was not empty; identical on every CTA; the barrier also
				                                            // puts every thread's Adam stores in front of the state's publication)
				if (b + 1 < tp.num_batches)
					pending_state = st, publish_pending = true; // every CTA has read the old state once the next barrier is passed
				else
					publish_state_if_last(adam, st);
			}
		}
		NRC_GTRACE(0x54);
		if (b + 1 < tp.num_batches) {
			grid_sync<false>(tp.grid_bar, bar_target); // the next batch runs on the weights (and optimizer state) just written
			NRC_GTRACE(0x59);
			if (publish_pending && blockIdx.x == 0 && threadIdx.x == 0)
				*tp.adam.opt_state = pending_state; // read again only after the next batch's first grid barrier
		}
		n = n_next, my_tiles = tiles_next;
	}
	NRC_GTRACE(10);
	} // warps 0..8
#ifdef NRC_TRACE
	if (blockIdx.x == 0 && threadIdx.x == 0)
		g_nrc_gtrace_n = gtrace_n;
	if (blockIdx.x == 0 && warp == kIssueWarp && itrace_n) // (only the elected issuing thread has logged anything)
		g_nrc_itrace_n = itrace_n;
	if (blockIdx.x == 0 && threadIdx.x == kProducerWarp0 * 32)
		g_nrc_ptrace_n = ptrace_n;
#endif
	tc_fence_before();
	__syncthreads();
	if (warp == kIssueWarp)
		tmem_dealloc(tmem, 512);
}

uint32_t gradient_max_partials(int sms) { return (uint32_t)sms; }

template <int IN_MODE, bool MULTI>
static cudaError_t launch_train_t(const TrainParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, uint32_t grid, cudaStream_t stream) {
	auto kern = nrc_train_kernel<IN_MODE, MULTI>;
	static std::atomic<uint64_t> configured{0}; // function attributes are per device: one bit per device ordinal
	int dev = 0;
	if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess)
		return e;
	const uint64_t dev_bit = 1ull << (dev & 63);
	if (!(configured.load(std::memory_order_acquire) & dev_bit)) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, train_smem_bytes(kPoolMulti));
		if (e != cudaSuccess)
			return e;
		configured.fetch_or(dev_bit, std::memory_order_release);
	}
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kTrainThreads), cfg.dynamicSmemBytes = train_smem_bytes(p.pool_tiles), cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeCooperative; // the grid barrier needs every CTA resident
	attr[0].val.cooperative = 1;
	cfg.attrs = attr, cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, p, tm_w, tm_in);
}

cudaError_t launch_train(const TrainParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream, uint32_t *grid_out) {
	uint64_t ntiles = 1;
	for (uint32_t b = 0; b < p.num_batches; ++b) {
		const uint64_t t = (p.batch[b].n + NRC_TILE - 1) / NRC_TILE;
		ntiles = t > ntiles ? t : ntiles;
	}
	// One CTA per tile up to the number of SMs - but never fewer than 108: the reduction / exchange / optimizer phase handles
	// the 324 blocks of the gradient three per CTA and round, and with a small batch (a shard of 2048 records is 16 tiles) a
	// 16-CTA grid would walk seven rounds one after the other, every one with its own wait for the peers' words (8 GPUs, 4 x
	// 16 384 records sharded: 208 us per frame). CTAs without a tile skip the gradient phase and write no partial.
	const uint64_t min_grid = 108 < sms ? 108 : sms;
	const uint64_t want = ntiles > min_grid ? ntiles : min_grid;
	const uint32_t grid = (uint32_t)(want < (uint64_t)sms ? want : (uint64_t)sms);
	if (grid_out)
		*grid_out = grid;
	TrainParams q = p;
	q.pool_tiles = ntiles > grid ? kPoolMulti : kPoolSingle; // two tiles in flight (+ a spare input buffer) only where a CTA has more than one
	const bool multi = p.comm.world > 1;
	switch (p.batch[0].in_mode) {
	case NRC_IN_ENCODED:
		return multi ? launch_train_t<NRC_IN_ENCODED, true>(q, tm_w, tm_in, grid, stream) : launch_train_t<NRC_IN_ENCODED, false>(q, tm_w, tm_in, grid, stream);
	case NRC_IN_UNPACKED:
		return multi ? launch_train_t<NRC_IN_UNPACKED, true>(q, tm_w, tm_in, grid, stream) : launch_train_t<NRC_IN_UNPACKED, false>(q, tm_w, tm_in, grid, stream);
	case NRC_IN_IMAGE_RANDOM:
		return launch_train_t<NRC_IN_IMAGE_RANDOM, false>(q, tm_w, tm_in, grid, stream); // (the learn-an-image harness is single-GPU)
	case NRC_IN_PACKED:
		return multi ? launch_train_t<NRC_IN_PACKED, true>(q, tm_w, tm_in, grid, stream) : launch_train_t<NRC_IN_PACKED, false>(q, tm_w, tm_in, grid, stream);
	}
	return cudaErrorInvalidValue;
}

// stand-alone optimizer step (used when an all-reduce sits between the reduction and the step)
__global__ void __launch_bounds__(128) adam_kernel(const AdamParams a) {
	const float count = a.gradients[NRC_GRAD_COUNT_SLOT];
	if (!(count > 0.0f)) // nrc_optimize.comp:33-34 / nrc_train_prepare.comp:22
		return;
	const NrcOptimizerState st = advance_state(*a.opt_state);
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < NRC_WEIGHT_COUNT)
		adam_update(a, i, a.entries[i], a.gradients[i], count, st);
	__syncthreads();
	publish_state_if_last(a, st);
}

cudaError_t launch_adam(const AdamParams &a, cudaStream_t stream) {
	adam_kernel<<<(NRC_WEIGHT_COUNT + 127) / 128, 128, 0, stream>>>(a);
	return cudaGetLastError();
}

// mlp_learning_an_image/optimize.comp:21-29
__global__ void __launch_bounds__(128) sgd_kernel(const SgdParams p) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= NRC_WEIGHT_COUNT)
		return;
	const float g = __fdiv_rn(__fdiv_rn(p.gradients[i], p.batch), 1.0f);
	if (isnan(g) || isinf(g))
		return;
	const float w = __fsub_rn(p.entries[i].weight, __fmul_rn(p.lr, g));
	p.entries[i].weight = w;
	p.weights[i] = __float2half_rn(w);
}
cudaError_t launch_sgd(const SgdParams &p, cudaStream_t stream) {
	sgd_kernel<<<(NRC_WEIGHT_COUNT + 127) / 128, 128, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace nrc
