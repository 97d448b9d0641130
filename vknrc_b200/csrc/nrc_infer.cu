// nrc_infer.cu -- query inference of the NRC MLP on sm_100a (replaces shader/src/nrc_inference.comp:30-74,
// test/evaluate_NV.comp:16-31 and test/mlp_learning_an_image/inference.comp:32-54 of the reference).
//
// One persistent CTA per SM, NT "slots" of 4 warps, no other warps. Per CTA:
//   * all six weight matrices are staged once (TMA tensor loads with the 128-byte swizzle straight from the
//     reference's row-major fp16 buffer; rows past 323 are zero-filled by TMA, which pads W5 from 3 to 64 rows for free);
//   * each slot owns one 128-sample tile in flight (64 accumulator + 32 operand TMEM columns). Layer l of a slot is
//     D[128 samples x 64] (fp32, TMEM) = A_l (fp16, TMEM; pre-encoded layer 0 straight from the slot's TMA buffer) x
//     W_l^T, issued as 4 tcgen05.mma (K=16) by one elected thread of one of the slot's warps.
//     Epilogue: thread = sample = TMEM lane: tcgen05.ld the accumulator row, ReLU + fp32->fp16 in one
//     cvt.rn.relu.f16x2.f32 per pair, tcgen05.st it back as the next layer's A operand. Activations never leave
//     TMEM. For record inputs the same threads run the input encoding and write A_0 directly.
//   * every slot double-buffers its own 16 KB input tiles (A_0, 128-byte-swizzled rows) in shared memory; layer 0 reads
//     them SS-form. Pre-encoded inputs: the slot's issuing thread TMA-loads the tile after next as soon as the layer-0
//     MMA that read a buffer has completed. Record inputs (14-float records, 16-byte packed records + scene gather,
//     image grid): NP extra PRODUCER warps unpack + encode records ahead of the slots (a warp = a quarter tile at a
//     time, up to two tiles ahead per slot) and write the encoded rows into the same buffers - the input encoding is
//     fused into the first layer, and the gather / encode latency never sits in a slot's MMA -> epilogue chain.
// Hand-offs: a 128-thread named barrier inside the slot (operand stored / accumulator drained -> issuer), the
// mbarriers d_full[slot] (tcgen05.commit -> epilogue), in_full[slot][2] (TMA or 4 producer warps -> layer-0 MMA) and
// in_free[slot][2] (a use counter per buffer: layer 0 done -> producers).
#include "nrc_kernels.h"
#include <atomic>
#include "nrc_encode.cuh"
#include "nrc_unpack.cuh"

using namespace sm100;

#ifdef NRC_TRACE
// development aid (tools/trace_infer.cu): thread 0 of each slot of CTA 0 logs (tag, clock) pairs into shared memory,
// dumped to global memory when the kernel ends; tag = event<<24 | layer<<16 | tile. Costs a few cycles per event.
#define NRC_TRACE_CAP 512
__device__ uint2 g_nrc_trace[4][NRC_TRACE_CAP];
__device__ unsigned int g_nrc_trace_n[4];
#define NRC_TRACE_EV(who, tag)                                                                                         \
	do {                                                                                                               \
		if (blockIdx.x == 0 && (threadIdx.x & 127) == 0 && (who) < 4 && trace_n < NRC_TRACE_CAP)                       \
			trace_sm[(who) * NRC_TRACE_CAP + trace_n++] = make_uint2((uint32_t)(tag), (uint32_t)clock64());             \
	} while (0)
#else
#define NRC_TRACE_EV(who, tag)
#endif
#define NRC_TRACE_TAG(ev) ((uint32_t)(ev) << 24 | (uint32_t)l << 16 | (j & 0xffffu))

namespace nrc {

// Shared memory: weights (6 x 8 KB), two 16 KB input buffers per slot, barriers.
constexpr uint32_t kSmemTextures = 64; // texture descriptors staged in shared memory when the scene has no more than this
template <int NT, int IN_MODE> struct InferSmem {
	static constexpr uint32_t kWeightBytes = NRC_LAYERS * 8192;
	static constexpr uint32_t kInOff = kWeightBytes;
	static constexpr uint32_t kInBytes = NT * 2 * 16384;
	static constexpr uint32_t kBarOff = kInOff + kInBytes;
	static constexpr uint32_t kLutOff = kBarOff + 256; // sRGB -> linear table (packed records only)
	static constexpr uint32_t kLutBytes = IN_MODE == NRC_IN_PACKED ? 1024 + kSmemTextures * 16 : 0; // + texture descriptors
#ifdef NRC_TRACE
	static constexpr uint32_t kTraceOff = kLutOff + kLutBytes;
	static constexpr uint32_t kBytes = kTraceOff + 4 * NRC_TRACE_CAP * 8 + 1024;
#else
	static constexpr uint32_t kBytes = kLutOff + kLutBytes + 1024; // + slack for manual 1024 B alignment
#endif
};

__device__ __forceinline__ uint32_t pack_rgba8(float r, float g, float b) { // imageStore to rgba8 (inference.comp:53)
	auto q = [](float x) { return (uint32_t)__float2int_rn(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f); };
	return q(r) | (q(g) << 8) | (q(b) << 16) | (255u << 24);
}

// Screen-destined results composite into two images read-modify-write (nrc_inference.comp:53-59). The slot threads fetch
// the dst word and then the pixel's bias / factor a few layers ahead of the write (ScatterPrefetch), so that neither
// dependent load level sits in front of the slot's next MMA.
struct ScatterPrefetch {
	uint32_t dst;
	float4 bf;
	float2 gb;
};
__device__ __forceinline__ void prefetch_dst(const InferParams &p, uint64_t gi, bool valid, ScatterPrefetch &pf) {
	pf.dst = valid ? __ldg(p.dst + gi * p.dst_stride_u32) : NRC_EVAL_INVALID_DST;
}
__device__ __forceinline__ void prefetch_pixel(const InferParams &p, ScatterPrefetch &pf) {
	if (pf.dst != NRC_EVAL_INVALID_DST && (pf.dst & 1u) == 0u) {
		const uint32_t e = pf.dst >> 1, x = e & 0x7FFFu, y = e >> 15;
		const uint64_t at = (uint64_t)y * p.image_pitch + x;
		pf.bf = *((const float4 *)p.bias_factor_r + at);
		pf.gb = ((const float2 *)p.factor_gb)[at];
	}
}

// nrc_inference.comp:48-73
__device__ __forceinline__ void write_result(const InferParams &p, uint64_t gi, float y0, float y1, float y2, const ScatterPrefetch &pf) {
	if (p.out_mode == NRC_OUT_F16VEC3) {
		if (p.clamp_output)
			y0 = fmaxf(y0, 0.0f), y1 = fmaxf(y1, 0.0f), y2 = fmaxf(y2, 0.0f);
		__half *o = (__half *)p.out + 3 * gi;
		o[0] = __float2half_rn(y0), o[1] = __float2half_rn(y1), o[2] = __float2half_rn(y2);
	} else if (p.out_mode == NRC_OUT_RGBA8) {
		((uint32_t *)p.out)[gi] = pack_rgba8(y0, y1, y2);
	} else { // NRC_OUT_SCATTER
		// NNOutput3 hands the prediction over as fp16 widened to fp32 (NN_nv.glsl:148-158), then max(., 0) (nrc_inference.comp:48)
		y0 = fmaxf(__half2float(__float2half_rn(y0)), 0.0f), y1 = fmaxf(__half2float(__float2half_rn(y1)), 0.0f),
		y2 = fmaxf(__half2float(__float2half_rn(y2)), 0.0f);
		const uint32_t dst = pf.dst;
		if (dst == NRC_EVAL_INVALID_DST)
			return;
		if ((dst & 1u) == 0u) {
			const uint32_t e = dst >> 1, x = e & 0x7FFFu, y = e >> 15;
			float4 *bf = (float4 *)p.bias_factor_r + (uint64_t)y * p.image_pitch + x;
			const float4 v = pf.bf;
			*bf = make_float4(v.x + v.w * y0, v.y + pf.gb.x * y1, v.z + pf.gb.y * y2, 0.0f);
		} else {
			const uint32_t e = dst >> 1, b = e & 3u, l = (e >> 2) & 0x3FFFu, r = e >> 16;
			NrcTrainRecord *recs = (NrcTrainRecord *)p.train_records[b];
			for (uint32_t i = l; i <= r; ++i) {
				NrcTrainRecord *t = recs + i;
				t->bias_r = t->bias_r + t->factor_r * y0;
				t->bias_g = t->bias_g + t->factor_g * y1;
				t->bias_b = t->bias_b + t->factor_b * y2;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ producer warps
// One unit of producer work = a quarter tile (32 records, one per lane): raw record -> 64 encoded features -> the
// row's eight 16-byte chunks at their 128-byte-swizzle positions (exactly what TMA would have written).
template <int IN_MODE> struct RawInput {
	uint32_t pk[4]; // NRC_IN_PACKED
	float2 f[7];    // NRC_IN_UNPACKED
};
template <int IN_MODE> __device__ __forceinline__ void load_raw(const InferParams &p, uint64_t gi, bool valid, RawInput<IN_MODE> &r) {
	if (IN_MODE == NRC_IN_PACKED) {
		r.pk[0] = r.pk[1] = r.pk[2] = r.pk[3] = 0u;
		if (valid)
			load_packed_input(p.in, gi, p.in_stride_bytes, r.pk);
	} else if (IN_MODE == NRC_IN_UNPACKED) {
		const float2 *src = (const float2 *)((const uint8_t *)p.in + gi * p.in_stride_bytes);
#pragma unroll
		for (int i = 0; i < 7; ++i)
			r.f[i] = valid ? __ldg(src + i) : make_float2(0.0f, 0.0f);
	}
}
template <int IN_MODE>
__device__ __forceinline__ void encode_raw(const InferParams &p, uint64_t gi, bool valid, const RawInput<IN_MODE> &r, uint32_t o[32], const float *lut,
                                           const NrcTexture *textures) {
	if (IN_MODE == NRC_IN_IMAGE_GRID) { // uv = (coord + 0.5) / width  (inference.comp:33-34)
		const uint32_t x = (uint32_t)(gi % p.image_width), y = (uint32_t)(gi / p.image_width);
		encode_oneblob32(((float)x + 0.5f) / (float)p.image_width, ((float)y + 0.5f) / (float)p.image_width, o);
		return;
	}
	float in[14];
	if (IN_MODE == NRC_IN_PACKED) {
		if (valid) {
			unpack_nrc_input(p.scene, r.pk, in, lut, textures);
		} else {
#pragma unroll
			for (int i = 0; i < 14; ++i)
				in[i] = 0.0f;
		}
	} else {
#pragma unroll
		for (int i = 0; i < 7; ++i)
			in[2 * i] = r.f[i].x, in[2 * i + 1] = r.f[i].y;
	}
	encode_nrc(in, o);
}

template <int NT, int NP, int IN_MODE>
__global__ void __launch_bounds__((NT * 4 + NP) * 32, 1)
    nrc_infer_kernel(const InferParams p, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in) {
	using L = InferSmem<NT, IN_MODE>;
	static_assert((IN_MODE == NRC_IN_ENCODED) == (NP == 0), "producer warps exist exactly for the record input modes");
	extern __shared__ uint8_t smem_raw[];
	uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *w_sm = smem;
	uint64_t *bars = (uint64_t *)(smem + L::kBarOff);
	uint64_t *w_full = bars, *d_full = bars + 1, *in_full = d_full + NT; // in_full[slot][2]
	volatile uint32_t *in_free = (volatile uint32_t *)(in_full + 2 * NT); // in_free[slot][2]: uses of the buffer the slot is done with
	uint32_t *tmem_slot = (uint32_t *)(in_free + 2 * NT);
#ifdef NRC_TRACE
	uint2 *trace_sm = (uint2 *)(smem + L::kTraceOff);
	uint32_t trace_n = 0;
#endif

	// warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it (slot,
	// TMEM addresses, UMMA descriptors) in uniform registers - no R2UR chains in front of the tcgen05 instructions
	const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
#ifndef NRC_INFER_NO_PDL
	// Programmatic dependent launch: the next kernel of the stream may place its CTAs on an SM the moment this kernel's
	// CTA there exits (every CTA needs the whole TMEM and most of the shared memory, so nothing overlaps - what goes away
	// is the grid-completion -> launch latency between back-to-back launches: -1.0 .. -1.9 us per 1080p launch).
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
	// prologue that depends on nothing an earlier kernel wrote: barriers, TMEM
	if (threadIdx.x == 0) {
		mbar_init(w_full, 1);
		for (int i = 0; i < NT; ++i) {
			mbar_init(d_full + i, 1);
			for (int b = 0; b < 2; ++b)
				mbar_init(in_full + 2 * i + b, NP ? 4 : 1), in_free[2 * i + b] = 0u;
		}
		fence_mbar_init();
		tma_prefetch_desc(&tm_w);
	}
	if (warp == 0)
		tmem_alloc(tmem_slot, 512);
#ifndef NRC_INFER_NO_PDL
	// everything the previous kernel of the stream wrote (records, counts, weights, scene tables) is read after this point
	asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
	uint64_t n = p.n;
	if (p.d_count) { // device-resident count, like the reference's indirect dispatch (nrc_indirect.comp:10)
		const uint64_t c = *p.d_count;
		n = c < n ? c : n;
	}
	const uint32_t ntiles = (uint32_t)((n + NRC_TILE - 1) / NRC_TILE);
	const uint32_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

	const float *lut = (const float *)(smem + L::kLutOff);
	const NrcTexture *textures = p.scene.textures;
	if (IN_MODE == NRC_IN_PACKED) {
		if (threadIdx.x < 256)
			((float *)(smem + L::kLutOff))[threadIdx.x] = kSrgbToLinear[threadIdx.x];
		if (p.scene.texture_count <= kSmemTextures) {
			NrcTexture *tsm = (NrcTexture *)(smem + L::kLutOff + 1024);
			if (threadIdx.x < p.scene.texture_count)
				tsm[threadIdx.x] = p.scene.textures[threadIdx.x];
			textures = tsm;
		}
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
	if (my_tiles == 0) { // (the device-resident count left this CTA without work)
		if (warp == 0)
			tmem_dealloc(tmem, 512);
		return;
	}
	if (threadIdx.x == 0) { // all six weight matrices, once per CTA
		mbar_arrive_expect_tx(w_full, L::kWeightBytes);
		for (int l = 0; l < NRC_LAYERS; ++l)
			tma_load_2d(w_sm + l * 8192, &tm_w, 0, l * 64, w_full);
	}

	if (NP > 0 && warp >= NT * 4) {
		// ------------------------------------------------------------------------------------------ producer warps
		// unit u = quarter (u & 3) of the CTA's tile g = u >> 2, which slot g % NT runs as its (g / NT)-th tile from
		// buffer (g / NT) & 1. A buffer is handed back through a use COUNTER in shared memory (the slot's issuing thread
		// stores "uses finished"; the producer of use m waits for the counter to reach m): unlike a parity wait this cannot
		// miss a phase, so any number of producer warps may take the units round-robin. The raw record of the warp's next
		// unit is fetched before the current one is processed.
		const uint32_t pw = warp - NT * 4, units = 4 * my_tiles;
		auto unit_gi = [&](uint32_t u) { return (uint64_t)(blockIdx.x + (u >> 2) * gridDim.x) * NRC_TILE + (u & 3u) * 32 + lane; };
		RawInput<IN_MODE> raw, raw_next;
		if (pw < units)
			load_raw<IN_MODE>(p, unit_gi(pw), unit_gi(pw) < n, raw);
#pragma unroll 1
		for (uint32_t u = pw; u < units; u += NP) {
			const uint32_t g = u >> 2, s = g % NT, it = g / NT, b = it & 1u, row = (u & 3u) * 32 + lane;
			const uint64_t gi = unit_gi(u);
			if (u + NP < units)
				load_raw<IN_MODE>(p, unit_gi(u + NP), unit_gi(u + NP) < n, raw_next);
			uint32_t o[32];
			encode_raw<IN_MODE>(p, gi, gi < n, raw, o, lut, textures);
			if (it >= 2) { // the slot has finished layer 0 of the tile that used this buffer before
				for (uint32_t spins = 0; ld_acquire_cta(in_free + 2 * s + b) < (it >> 1); ++spins) {
					if (spins > (1u << 22))
						__trap(); // a protocol bug must not hang the GPU
					__nanosleep(NRC_INFER_FREE_BACKOFF);
				}
				__threadfence_block();
			}
			uint8_t *dst = smem + L::kInOff + (s * 2 + b) * 16384 + row * 128;
#pragma unroll
			for (int c = 0; c < 8; ++c)
				*(uint4 *)(dst + ((c ^ (row & 7)) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
			fence_proxy_async_smem(); // generic-proxy stores -> visible to the MMA's async-proxy operand reads
			__syncwarp();
			if (lane == 0)
				mbar_arrive(in_full + 2 * s + b);
			raw = raw_next;
		}
	} else {
	// ---------------------------------------------------------------------------------------------- slot warpgroups
	// Each slot is self-contained: its 128 threads run the epilogues, and one elected thread of ONE of its warps issues
	// the slot's tcgen05.mma (and, for pre-encoded inputs, the TMA loads of the slot's own double-buffered input tiles).
	// The issuing warp rotates with the slot index so that the issue work spreads over the four SM sub-partitions.
	const uint32_t s = warp >> 2, q = warp & 3, row = q * 32 + lane;
	const bool issuer_warp = q == (s & 3u);
	const uint32_t d_col = tmem + s * 96, a_col = d_col + 64;             // issuer's view (lane 0)
	const uint32_t d_t = tmem_addr(tmem, q * 32, s * 96), a_t = d_t + 64; // this warp's 32 lanes
	uint8_t *in_sm = smem + L::kInOff + s * 2 * 16384;
	uint64_t *my_in_full = in_full + 2 * s, *my_d_full = d_full + s;
#ifdef NRC_INFER_F16ACC
	constexpr uint32_t idesc64 = make_idesc_f16_f16(128, 64, false, false);
	constexpr uint32_t idesc16 = make_idesc_f16_f16(128, 16, false, false);
#else
	constexpr uint32_t idesc64 = make_idesc_f16_f32(128, 64, false, false);
	constexpr uint32_t idesc16 = make_idesc_f16_f32(128, 16, false, false);
#endif
	// UMMA descriptors differ only in the start-address field: desc(addr + off) = desc(addr) + (off >> 4)
	// (only the start-address field, i.e. the low word, varies: see mma_*_lh)
	const uint32_t w_lo = smem_desc_lo(smem_u32(w_sm)), in_lo = smem_desc_lo(smem_u32(in_sm));
	auto slot_sync = [&]() { // all of this slot's TMEM traffic is complete and visible to the issuer
		tc_fence_before();
		asm volatile("bar.sync %0, 128;" ::"r"(s + 1) : "memory");
	};
	auto load_input_tile = [&](uint32_t it) { // issuer thread only: tile `it` of this slot -> buffer it & 1
		const uint32_t tile = blockIdx.x + (s + it * NT) * gridDim.x;
		mbar_arrive_expect_tx(my_in_full + (it & 1), 16384);
		tma_load_2d(in_sm + (it & 1) * 16384, &tm_in, 0, (int32_t)(tile * NRC_TILE), my_in_full + (it & 1));
	};
	const uint32_t slot_tiles = s < my_tiles ? (my_tiles - s + NT - 1) / NT : 0;
	if (issuer_warp) {
		if (elect_one()) {
			if (IN_MODE == NRC_IN_ENCODED) {
				tma_prefetch_desc(&tm_in);
				for (uint32_t it = 0; it < 2 && it < slot_tiles; ++it)
					load_input_tile(it);
			}
			mbar_wait(w_full, 0);
		}
		__syncwarp();
	}
	uint32_t d_cnt = 0;
	// The result of a tile is written out only after the next tile's first layer has been issued: the global stores (and
	// the scatter's read-modify-write) then overlap that layer's MMAs instead of sitting in front of them.
	float pend_y0 = 0.0f, pend_y1 = 0.0f, pend_y2 = 0.0f;
	uint64_t pend_gi = 0;
	bool pend = false;
	ScatterPrefetch pf{};
	for (uint32_t it = 0; it < slot_tiles; ++it) {
		const uint32_t j = s + it * NT;
		const uint32_t tile = blockIdx.x + j * gridDim.x;
		const uint64_t gi = (uint64_t)tile * NRC_TILE + row;
		const bool valid = gi < n;
		uint32_t b_lo = w_lo; // descriptor (low word) of W_l, advanced by one 8 KB matrix per layer
#pragma unroll // layer index static: no per-layer branches, descriptors are immediates (70 -> 59.5 us at 1080p)
		for (int l = 0; l < NRC_LAYERS; ++l, b_lo += 8192 >> 4) {
			// ---- issue layer l of this slot's tile
			NRC_TRACE_EV(s, NRC_TRACE_TAG(1));
			if (issuer_warp) {
				if (elect_one()) {
					tc_fence_after();
					NRC_TRACE_EV(s, NRC_TRACE_TAG(6));
					if (l == 0) { // A_0 from the slot's input buffer (TMA or producer warps), SS form
						mbar_wait(my_in_full + (it & 1), (it >> 1) & 1);
						const uint32_t a_lo = in_lo + (it & 1) * (16384 >> 4);
#pragma unroll
						for (int k = 0; k < 4; ++k)
							mma_ss_lh(d_col, a_lo + k * 2, b_lo + k * 2, kSmemDescHiSw128, idesc64, k > 0);
					} else {
#pragma unroll
						for (int k = 0; k < 4; ++k)
							mma_ts_lh(d_col, a_col + k * 8, b_lo + k * 2, kSmemDescHiSw128, l < NRC_HIDDEN_LAYERS ? idesc64 : idesc16, k > 0);
					}
					NRC_TRACE_EV(s, NRC_TRACE_TAG(7));
					tc_commit(my_d_full);
					// layer 0 has consumed input buffer (it & 1) (its completion was observed below, one layer ago)
					if (l == 1) {
						if (IN_MODE == NRC_IN_ENCODED) {
							if (it + 2 < slot_tiles)
								load_input_tile(it + 2);
						} else {
							st_release_cta(in_free + 2 * s + (it & 1), (it >> 1) + 1); // (this thread observed layer 0's completion one layer ago)
						}
					}
					NRC_TRACE_EV(s, NRC_TRACE_TAG(0));
				}
				__syncwarp();
			}
			if (l == 0 && pend) {
				write_result(p, pend_gi, pend_y0, pend_y1, pend_y2, pf);
				pend = false;
			}
			if (p.out_mode == NRC_OUT_SCATTER) { // (after the previous tile's flush, which consumed the old prefetch)
				if (l == 1)
					prefetch_dst(p, gi, valid, pf);
				if (l == 3)
					prefetch_pixel(p, pf);
			}
			// ---- epilogue of layer l
			NRC_TRACE_EV(s, NRC_TRACE_TAG(2));
			mbar_wait(my_d_full, d_cnt & 1);
			++d_cnt;
			tc_fence_after();
			NRC_TRACE_EV(s, NRC_TRACE_TAG(3));
#ifdef NRC_INFER_F16ACC
			if (l < NRC_HIDDEN_LAYERS) {
				uint32_t v[32];
				tmem_ld_x32_pack16(d_t, v);
				tc_wait_ld();
#pragma unroll
				for (int i = 0; i < 32; ++i)
					v[i] = relu_f16x2(v[i]);
				NRC_TRACE_EV(s, NRC_TRACE_TAG(4));
				tmem_st_x32(a_t, v);
				tc_wait_st();
				NRC_TRACE_EV(s, NRC_TRACE_TAG(5));
				slot_sync();
			} else {
				uint32_t y[4];
				tmem_ld_x4(d_t, y);
				tc_wait_ld();
				slot_sync();
				pend = valid, pend_gi = gi;
				pend_y0 = __half2float(__ushort_as_half((unsigned short)y[0])), pend_y1 = __half2float(__ushort_as_half((unsigned short)y[1]));
				pend_y2 = __half2float(__ushort_as_half((unsigned short)y[2]));
			}
#else
			if (l < NRC_HIDDEN_LAYERS) {
				if (NP > 0) { // (1024-thread CTA, 64 registers: the row in two halves)
#pragma unroll
					for (int hh = 0; hh < 2; ++hh) {
						uint32_t v[32], o[16];
						tmem_ld_x32(d_t + 32 * hh, v);
						tc_wait_ld();
#pragma unroll
						for (int i = 0; i < 16; ++i)
							o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
						tmem_st_x16(a_t + 16 * hh, o);
					}
				} else {
					uint32_t v[64], o[32];
					tmem_ld_x32(d_t, v);
					tmem_ld_x32(d_t + 32, v + 32);
					tc_wait_ld();
#pragma unroll
					for (int i = 0; i < 16; ++i)
						o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
					tmem_st_x16(a_t, o);
#pragma unroll
					for (int i = 16; i < 32; ++i)
						o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
					NRC_TRACE_EV(s, NRC_TRACE_TAG(4));
					tmem_st_x16(a_t + 16, o + 16);
				}
				tc_wait_st();
				NRC_TRACE_EV(s, NRC_TRACE_TAG(5));
				slot_sync();
			} else {
				uint32_t y[4];
				tmem_ld_x4(d_t, y);
				tc_wait_ld();
				slot_sync(); // accumulator drained before the next tile's layer 0 overwrites it
				pend = valid, pend_gi = gi;
				pend_y0 = __uint_as_float(y[0]), pend_y1 = __uint_as_float(y[1]), pend_y2 = __uint_as_float(y[2]);
			}
#endif
		}
	}
	if (pend)
		write_result(p, pend_gi, pend_y0, pend_y1, pend_y2, pf);
#ifdef NRC_TRACE
	if (blockIdx.x == 0 && (threadIdx.x & 127) == 0 && (threadIdx.x >> 7) < 4) {
		for (uint32_t i = 0; i < trace_n; ++i)
			g_nrc_trace[threadIdx.x >> 7][i] = trace_sm[(threadIdx.x >> 7) * NRC_TRACE_CAP + i];
		g_nrc_trace_n[threadIdx.x >> 7] = trace_n;
	}
#endif
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0)
		tmem_dealloc(tmem, 512);
}

template <int NT, int NP, int IN_MODE> static cudaError_t launch(const InferParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream) {
	auto kern = nrc_infer_kernel<NT, NP, IN_MODE>;
	constexpr uint32_t smem_bytes = InferSmem<NT, IN_MODE>::kBytes;
	static std::atomic<uint64_t> configured{0}; // function attributes are per device: one bit per device ordinal
	int dev = 0;
	if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess)
		return e;
	const uint64_t dev_bit = 1ull << (dev & 63);
	if (!(configured.load(std::memory_order_acquire) & dev_bit)) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
		if (e != cudaSuccess)
			return e;
		configured.fetch_or(dev_bit, std::memory_order_release);
	}
	const uint64_t ntiles = (p.n + NRC_TILE - 1) / NRC_TILE;
	const uint32_t grid = (uint32_t)(ntiles < (uint64_t)sms ? ntiles : (uint64_t)sms);
#ifdef NRC_INFER_NO_PDL
	kern<<<grid, (NT * 4 + NP) * 32, smem_bytes, stream>>>(p, tm_w, tm_in);
	return cudaGetLastError();
#else
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid), cfg.blockDim = dim3((NT * 4 + NP) * 32), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; // see griddepcontrol.* at the top of the kernel
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr, cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, p, tm_w, tm_in);
#endif
}

// Stand-alone UnpackNRCInput (NRCRecord.glsl:98-125): [n] PackedNRCInput -> [n][14] fp32. The fused kernels above never
// materialise this; it exists for parity checks of the gather and as the HBM-bound record-streaming stage on its own.
__global__ void __launch_bounds__(256) nrc_unpack_kernel(const void *packed, uint32_t stride_bytes, uint64_t n, const NrcScene scene, float *out14) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	uint32_t pk[4];
	float in[14];
	load_packed_input(packed, i, stride_bytes, pk);
	unpack_nrc_input(scene, pk, in);
	float2 *dst = (float2 *)(out14 + 14 * i);
#pragma unroll
	for (int k = 0; k < 7; ++k)
		dst[k] = make_float2(in[2 * k], in[2 * k + 1]);
}
// Stand-alone NRCInputEncode (NRCRecord.glsl:77-95), optionally behind UnpackNRCInput (:98-125): [n] records -> [n][64] fp16 in the
// layout test/evaluate_NV.comp reads. The fused kernels never materialise this either; it is the encode / record-streaming
// stage on its own - bound by HBM (56 or 16 B in, 128 B out per record; ~400 instructions per record hide under the traffic) -
// and the producer of pre-encoded inputs for nrc_infer_encoded / nrc_gradient_encoded. One thread per record; the 128 x 128-byte
// tile of a block is staged in shared memory (16-byte chunk c of row r at chunk position c ^ (r & 7): conflict-free both ways)
// so that every warp-wide global store covers 512 contiguous bytes. The encoder is the very device function of the fused paths:
// bit-identical features.
template <int IN_MODE>
__global__ void __launch_bounds__(128) nrc_encode_kernel(const void *in, uint32_t stride_bytes, uint64_t n, const NrcScene scene, uint4 *out) {
	__shared__ __align__(128) uint4 tile[128 * 8];
	const uint32_t t = threadIdx.x;
	const uint64_t base = (uint64_t)blockIdx.x * 128, i = base + t;
	uint32_t o[32];
#pragma unroll
	for (int k = 0; k < 32; ++k)
		o[k] = 0u;
	if (i < n) {
		float f[14];
		if (IN_MODE == NRC_IN_PACKED) {
			uint32_t pk[4];
			load_packed_input(in, i, stride_bytes, pk);
			unpack_nrc_input(scene, pk, f);
		} else {
			const float2 *src = (const float2 *)((const uint8_t *)in + i * stride_bytes);
#pragma unroll
			for (int k = 0; k < 7; ++k) {
				const float2 v = __ldg(src + k);
				f[2 * k] = v.x, f[2 * k + 1] = v.y;
			}
		}
		encode_nrc(f, o);
	}
#pragma unroll
	for (int c = 0; c < 8; ++c)
		tile[t * 8 + (c ^ (t & 7))] = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
	__syncthreads();
	const uint64_t rows = n - base < 128 ? n - base : 128; // rows of this tile that exist
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const uint32_t idx = k * 128 + t, r = idx >> 3, c = idx & 7u; // a warp: 4 rows x 8 chunks = 512 contiguous bytes
		if (r < rows)
			out[base * 8 + idx] = tile[r * 8 + (c ^ (r & 7))];
	}
}
cudaError_t launch_encode(const void *in, int in_mode, uint32_t stride_bytes, uint64_t n, const NrcScene &scene, void *out, cudaStream_t stream) {
	if (n == 0)
		return cudaSuccess;
	const unsigned grid = (unsigned)((n + 127) / 128);
	if (in_mode == NRC_IN_PACKED)
		nrc_encode_kernel<NRC_IN_PACKED><<<grid, 128, 0, stream>>>(in, stride_bytes, n, scene, (uint4 *)out);
	else
		nrc_encode_kernel<NRC_IN_UNPACKED><<<grid, 128, 0, stream>>>(in, stride_bytes, n, scene, (uint4 *)out);
	return cudaGetLastError();
}
// nrc_scene_build_prim_table: one thread per primitive copies what UnpackNRCInput would gather into its 64-byte row
__global__ void __launch_bounds__(256) nrc_prim_table_kernel(const NrcScene scene, uint32_t prim_count, NrcPrimRow *rows) {
	const uint32_t prim = blockIdx.x * blockDim.x + threadIdx.x;
	if (prim >= prim_count)
		return;
	NrcPrimRow r;
	for (int k = 0; k < 3; ++k) {
		const float *p = scene.vertices + 3 * (size_t)scene.vertex_indices[3 * (size_t)prim + k];
		r.v[k][0] = p[0], r.v[k][1] = p[1], r.v[k][2] = p[2];
		const float *t = scene.texcoords + 2 * (size_t)scene.texcoord_indices[3 * (size_t)prim + k];
		r.tc[k][0] = t[0], r.tc[k][1] = t[1];
	}
	r.material_id = scene.material_ids[prim];
	const float4 *src = (const float4 *)&r;
	float4 *dst = (float4 *)(rows + prim);
	for (int i = 0; i < 4; ++i)
		dst[i] = src[i];
}
cudaError_t launch_prim_table(const NrcScene &scene, uint32_t prim_count, void *rows, cudaStream_t stream) {
	if (prim_count == 0)
		return cudaSuccess;
	nrc_prim_table_kernel<<<(prim_count + 255) / 256, 256, 0, stream>>>(scene, prim_count, (NrcPrimRow *)rows);
	return cudaGetLastError();
}
cudaError_t launch_unpack(const void *packed, uint32_t stride_bytes, uint64_t n, const NrcScene &scene, float *out14, cudaStream_t stream) {
	if (n == 0)
		return cudaSuccess;
	nrc_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(packed, stride_bytes, n, scene, out14);
	return cudaGetLastError();
}

cudaError_t launch_infer(const InferParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream) {
	if (p.n == 0)
		return cudaSuccess;
	constexpr int NT = NRC_INFER_SLOTS, NTP = NRC_INFER_SLOTS_REC, NP = NRC_INFER_PRODUCER_WARPS;
	switch (p.in_mode) {
	case NRC_IN_ENCODED:
		return launch<NT, 0, NRC_IN_ENCODED>(p, tm_w, tm_in, sms, stream);
	case NRC_IN_UNPACKED:
		return launch<NTP, NP, NRC_IN_UNPACKED>(p, tm_w, tm_in, sms, stream);
	case NRC_IN_IMAGE_GRID:
		return launch<NTP, NP, NRC_IN_IMAGE_GRID>(p, tm_w, tm_in, sms, stream);
	case NRC_IN_PACKED:
		return launch<NRC_INFER_SLOTS_PACKED, NRC_INFER_PRODUCER_WARPS_PACKED, NRC_IN_PACKED>(p, tm_w, tm_in, sms, stream);
	}
	return cudaErrorInvalidValue;
}

} // namespace nrc
