// nrc_infer.cu -- query inference of the NRC MLP on sm_100a (replaces shader/src/nrc_inference.comp:30-74,
// test/evaluate_NV.comp:16-31 and test/mlp_learning_an_image/inference.comp:32-54 of the reference).
//
// One persistent CTA per SM. Per CTA:
//   * warp TMA  : stages all six weight matrices once (TMA tensor loads with the 128-byte swizzle straight from the
//                 reference's row-major fp16 buffer; rows past 323 are zero-filled by TMA, which pads W5 from 3 to 64
//                 rows for free) and, for pre-encoded inputs, streams 128x64 fp16 input tiles through a smem ring;
//   * NT "slots" of 4 warps, each owning one 128-sample tile in flight (64 accumulator + 32 operand TMEM columns).
//                 Layer l of a slot is D[128 samples x 64] (fp32, TMEM) = A_l (fp16, TMEM; pre-encoded layer 0 from the
//                 smem ring) x W_l^T, issued as 4 tcgen05.mma (K=16) by one elected thread of the slot's first warp.
//                 Epilogue: thread = sample = TMEM lane: tcgen05.ld the accumulator row, ReLU + fp32->fp16 in one
//                 cvt.rn.relu.f16x2.f32 per pair, tcgen05.st it back as the next layer's A operand. Activations never
//                 leave TMEM. For record inputs the same threads run the input encoding and write A_0 directly.
// Hand-offs: a 128-thread named barrier inside the slot (operand stored / accumulator drained -> issuer), the
// mbarrier d_full[slot] (tcgen05.commit -> epilogue) and in_full/in_empty[stage] (TMA <-> layer-0 MMA).
#include "nrc_kernels.h"
#include "nrc_encode.cuh"

using namespace sm100;

#ifdef NRC_TRACE
// development aid: CTA 0 records (tag, clock) pairs; tag = who<<24 | layer<<16 | tile
__device__ unsigned long long g_nrc_trace[4][4096];
__device__ unsigned int g_nrc_trace_n[4];
#define NRC_TRACE_EV(who, tag)                                                                                         \
	do {                                                                                                               \
		if (blockIdx.x == 0) {                                                                                         \
			unsigned int i_ = g_nrc_trace_n[who]++;                                                                    \
			if (i_ < 2048) {                                                                                           \
				g_nrc_trace[who][2 * i_] = (tag);                                                                      \
				g_nrc_trace[who][2 * i_ + 1] = clock64();                                                              \
			}                                                                                                          \
		}                                                                                                              \
	} while (0)
#else
#define NRC_TRACE_EV(who, tag)
#endif

namespace nrc {

template <int NT> struct InferSmem {
	static constexpr int kStages = 6;
	static constexpr uint32_t kWeightBytes = NRC_LAYERS * 8192;
	static constexpr uint32_t kRingOff = kWeightBytes;
	static constexpr uint32_t kBarOff = kRingOff + kStages * 16384;
	static constexpr uint32_t kBytes = kBarOff + 256 + 1024; // + slack for manual 1024 B alignment
};

__device__ __forceinline__ uint32_t pack_rgba8(float r, float g, float b) { // imageStore to rgba8 (inference.comp:53)
	auto q = [](float x) { return (uint32_t)__float2int_rn(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f); };
	return q(r) | (q(g) << 8) | (q(b) << 16) | (255u << 24);
}

// nrc_inference.comp:48-73
__device__ __forceinline__ void write_result(const InferParams &p, uint64_t gi, float y0, float y1, float y2) {
	if (p.out_mode == NRC_OUT_F16VEC3) {
		if (p.clamp_output)
			y0 = fmaxf(y0, 0.0f), y1 = fmaxf(y1, 0.0f), y2 = fmaxf(y2, 0.0f);
		__half *o = (__half *)p.out + 3 * gi;
		o[0] = __float2half_rn(y0), o[1] = __float2half_rn(y1), o[2] = __float2half_rn(y2);
	} else if (p.out_mode == NRC_OUT_RGBA8) {
		((uint32_t *)p.out)[gi] = pack_rgba8(y0, y1, y2);
	} else { // NRC_OUT_SCATTER
		y0 = fmaxf(y0, 0.0f), y1 = fmaxf(y1, 0.0f), y2 = fmaxf(y2, 0.0f);
		const uint32_t dst = p.dst[gi * p.dst_stride_u32];
		if (dst == NRC_EVAL_INVALID_DST)
			return;
		if ((dst & 1u) == 0u) {
			const uint32_t e = dst >> 1, x = e & 0x7FFFu, y = e >> 15;
			float4 *bf = (float4 *)p.bias_factor_r + (uint64_t)y * p.image_pitch + x;
			const float2 gb = ((const float2 *)p.factor_gb)[(uint64_t)y * p.image_pitch + x];
			float4 v = *bf;
			*bf = make_float4(v.x + v.w * y0, v.y + gb.x * y1, v.z + gb.y * y2, 0.0f);
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

template <int NT, int IN_MODE>
__global__ void __launch_bounds__((NT * 4 + 1) * 32, 1)
    nrc_infer_kernel(const InferParams p, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in) {
	using L = InferSmem<NT>;
	constexpr int kStages = L::kStages;
	constexpr uint32_t TMA_WARP = NT * 4;
	extern __shared__ uint8_t smem_raw[];
	uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *w_sm = smem;
	uint8_t *ring = smem + L::kRingOff;
	uint64_t *bars = (uint64_t *)(smem + L::kBarOff);
	uint64_t *w_full = bars, *in_full = bars + 1, *in_empty = in_full + kStages, *d_full = in_empty + kStages;
	uint32_t *tmem_slot = (uint32_t *)(d_full + NT);

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t n = p.n;
	if (p.d_count) { // device-resident count, like the reference's indirect dispatch (nrc_indirect.comp:10)
		const uint64_t c = *p.d_count;
		n = c < n ? c : n;
	}
	const uint32_t ntiles = (uint32_t)((n + NRC_TILE - 1) / NRC_TILE);
	if (blockIdx.x >= ntiles)
		return;
	const uint32_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

	if (threadIdx.x == 0) {
		mbar_init(w_full, 1);
		for (int i = 0; i < kStages; ++i)
			mbar_init(in_full + i, 1), mbar_init(in_empty + i, 1);
		for (int i = 0; i < NT; ++i)
			mbar_init(d_full + i, 1);
		fence_mbar_init();
	}
	if (warp == TMA_WARP)
		tmem_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *tmem_slot;

	if (warp == TMA_WARP) {
		// ------------------------------------------------------------------------------------------ TMA producer
		if (elect_one()) {
			tma_prefetch_desc(&tm_w);
			mbar_arrive_expect_tx(w_full, L::kWeightBytes);
			for (int l = 0; l < NRC_LAYERS; ++l)
				tma_load_2d(w_sm + l * 8192, &tm_w, 0, l * 64, w_full);
			if (IN_MODE == NRC_IN_ENCODED) {
				tma_prefetch_desc(&tm_in);
				for (uint32_t j = 0; j < my_tiles; ++j) {
					const uint32_t st = j % kStages, ph = (j / kStages) & 1;
					mbar_wait(in_empty + st, ph ^ 1);
					mbar_arrive_expect_tx(in_full + st, 16384);
					const uint32_t tile = blockIdx.x + j * gridDim.x;
					tma_load_2d(ring + st * 16384, &tm_in, 0, (int32_t)(tile * NRC_TILE), in_full + st);
				}
			}
		}
		__syncwarp();
	} else {
		// ------------------------------------------------------------------------------------------ slot warpgroups
		// Each slot is self-contained: its 128 threads run the epilogues and one elected thread of its first warp issues
		// the slot's own tcgen05.mma. A dedicated issuer thread for all slots serialises ~300 cycles of issue + wait
		// latency per layer-step (measured) and starves the tensor pipe; NT issuers in parallel do not.
		const uint32_t s = warp >> 2, q = warp & 3, row = q * 32 + lane;
		const uint32_t d_col = tmem + s * 96, a_col = d_col + 64;                        // issuer view (lane 0)
		const uint32_t d_t = tmem_addr(tmem, q * 32, s * 96), a_t = d_t + 64;            // this warp's 32 lanes
		const uint32_t w_addr = smem_u32(w_sm), ring_addr = smem_u32(ring);
		constexpr uint32_t idesc64 = make_idesc_f16_f32(128, 64, false, false);
		constexpr uint32_t idesc16 = make_idesc_f16_f32(128, 16, false, false);
		auto slot_sync = [&]() { // all of this slot's TMEM traffic is complete and visible to the issuer
			tc_fence_before();
			asm volatile("bar.sync %0, 128;" ::"r"(s + 1) : "memory");
		};
		uint32_t d_cnt = 0;
		if (q == 0)
			mbar_wait(w_full, 0);
		for (uint32_t j = s; j < my_tiles; j += NT) {
			const uint32_t tile = blockIdx.x + j * gridDim.x;
			const uint64_t gi = (uint64_t)tile * NRC_TILE + row;
			const bool valid = gi < n;
			if (IN_MODE != NRC_IN_ENCODED) {
				uint32_t o[32];
				if (IN_MODE == NRC_IN_UNPACKED) {
					float in[14];
					if (valid) {
						const float2 *src = (const float2 *)((const uint8_t *)p.in + gi * p.in_stride_bytes);
#pragma unroll
						for (int i = 0; i < 7; ++i) {
							const float2 t = __ldg(src + i);
							in[2 * i] = t.x, in[2 * i + 1] = t.y;
						}
					} else {
#pragma unroll
						for (int i = 0; i < 14; ++i)
							in[i] = 0.0f;
					}
					encode_nrc(in, o);
				} else { // NRC_IN_IMAGE_GRID: uv = (coord + 0.5) / width  (inference.comp:33-34)
					const uint32_t x = (uint32_t)(gi % p.image_width), y = (uint32_t)(gi / p.image_width);
					encode_oneblob32(((float)x + 0.5f) / (float)p.image_width, ((float)y + 0.5f) / (float)p.image_width, o);
				}
				tmem_st_x32(a_t, o);
				tc_wait_st();
				slot_sync();
			}
#pragma unroll 1
			for (int l = 0; l < NRC_LAYERS; ++l) {
				// ---- issue layer l of this slot's tile
				if (q == 0) {
					if (elect_one()) {
						tc_fence_after();
						const uint32_t b_addr = w_addr + l * 8192;
						if (IN_MODE == NRC_IN_ENCODED && l == 0) {
							const uint32_t st = j % kStages;
							mbar_wait(in_full + st, (j / kStages) & 1);
							const uint32_t a_addr = ring_addr + st * 16384;
#pragma unroll
							for (int k = 0; k < 4; ++k)
								mma_ss(d_col, make_smem_desc_sw128(a_addr + k * 32, 0, 1024), make_smem_desc_sw128(b_addr + k * 32, 0, 1024),
								       idesc64, k > 0);
							tc_commit(in_empty + st);
						} else {
#pragma unroll
							for (int k = 0; k < 4; ++k)
								mma_ts(d_col, a_col + k * 8, make_smem_desc_sw128(b_addr + k * 32, 0, 1024), l < NRC_HIDDEN_LAYERS ? idesc64 : idesc16,
								       k > 0);
						}
						tc_commit(d_full + s);
					}
					__syncwarp();
				}
				// ---- epilogue of layer l
				mbar_wait(d_full + s, d_cnt & 1);
				++d_cnt;
				tc_fence_after();
				if (l < NRC_HIDDEN_LAYERS) {
					uint32_t v[32], o[32];
					tmem_ld_x32(d_t, v);
					tc_wait_ld();
#pragma unroll
					for (int i = 0; i < 16; ++i)
						o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
					tmem_ld_x32(d_t + 32, v);
					tc_wait_ld();
#pragma unroll
					for (int i = 0; i < 16; ++i)
						o[16 + i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
					tmem_st_x32(a_t, o);
					tc_wait_st();
					slot_sync();
				} else {
					uint32_t y[4];
					tmem_ld_x4(d_t, y);
					tc_wait_ld();
					if (IN_MODE == NRC_IN_ENCODED)
						slot_sync(); // accumulator drained before the next tile's layer 0 overwrites it
					if (valid)
						write_result(p, gi, __uint_as_float(y[0]), __uint_as_float(y[1]), __uint_as_float(y[2]));
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == TMA_WARP)
		tmem_dealloc(tmem, 512);
}

template <int NT, int IN_MODE> static cudaError_t launch(const InferParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream) {
	auto kern = nrc_infer_kernel<NT, IN_MODE>;
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, InferSmem<NT>::kBytes);
		if (e != cudaSuccess)
			return e;
		configured = true;
	}
	const uint64_t ntiles = (p.n + NRC_TILE - 1) / NRC_TILE;
	const uint32_t grid = (uint32_t)(ntiles < (uint64_t)sms ? ntiles : (uint64_t)sms);
	kern<<<grid, (NT * 4 + 1) * 32, InferSmem<NT>::kBytes, stream>>>(p, tm_w, tm_in);
	return cudaGetLastError();
}

cudaError_t launch_infer(const InferParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, cudaStream_t stream) {
	if (p.n == 0)
		return cudaSuccess;
	constexpr int NT = NRC_INFER_SLOTS;
	switch (p.in_mode) {
	case NRC_IN_ENCODED:
		return launch<NT, NRC_IN_ENCODED>(p, tm_w, tm_in, sms, stream);
	case NRC_IN_UNPACKED:
		return launch<NT, NRC_IN_UNPACKED>(p, tm_w, tm_in, sms, stream);
	case NRC_IN_IMAGE_GRID:
		return launch<NT, NRC_IN_IMAGE_GRID>(p, tm_w, tm_in, sms, stream);
	}
	return cudaErrorInvalidValue;
}

} // namespace nrc
