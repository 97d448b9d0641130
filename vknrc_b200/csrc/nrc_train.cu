// nrc_train.cu -- online training of the NRC MLP on sm_100a: forward, loss gradient, back-propagation, batch
// weight-gradient reduction and the optimizer step. Replaces the reference's
//   shader/src/nrc_gradient.comp:26-58, test/train_NV.comp:18-46, test/mlp_learning_an_image/gradient.comp:46-80
//   (NN_nv.glsl: NNForward*, NNLoadDA3_*, NNBackwardDA*_ReLU, NNUpdateDW*),
//   shader/src/nrc_train_prepare.comp:16-28 + nrc_optimize.comp:32-54, mlp_learning_an_image/optimize.comp:21-29.
//
// Gradient kernel, per CTA (one 128-record tile at a time; 8 warps = 4 TMEM lane quarters x 2 column halves, one
// elected thread of warp 0 issues the TMA loads and every tcgen05.mma right after the CTA barrier):
//   every operand tile is an array of 128-byte rows (64 fp16) in shared memory with the 128-byte swizzle:
//     W_l      [out][in]     used K-major  (forward B)   and MN-major (dA: B with K = out)
//     a_l      [sample][in]  used K-major  (forward A)   and MN-major (dW: B with K = sample)
//     delta_l  [sample][out] used K-major  (dA: A)       and MN-major (dW: A with K = sample)
//   so each tensor is stored once and read through two UMMA descriptor flavours - no transposes, no copies.
//   forward  l=0..5 : D[128 x 64|16] = a_l * W_l^T                  (M=128)  -> ReLU -> a_{l+1}         (TMEM -> smem)
//   loss            : delta_5 = dL/dy (NN_nv.glsl:162-196)
//   backward l=5..1 : D[128 x 64]   = delta_l * W_l                 (M=128)  -> * [a_l > 0] -> delta_{l-1}
//   dW       l=5..0 : dW_l[64 x 64] += delta_l^T * a_l  (K = 128 samples, M=64) accumulated IN TMEM across all of the
//                     CTA's tiles (5*64 + 16 fp32 columns), written once per CTA as a partial.
// The partials are then summed in a fixed order by reduce_partials_kernel (deterministic, unlike the reference's
// 2.6 M fp32 atomics per batch, NN_nv.glsl:309-314,357-364), and adam_kernel applies nrc_optimize.comp verbatim.
#include "nrc_kernels.h"
#include "nrc_encode.cuh"

using namespace sm100;

namespace nrc {

namespace {
constexpr uint32_t kWOff = 0;                         // 6 x 8 KB weights
constexpr uint32_t kActOff = NRC_LAYERS * 8192;       // 6 x 16 KB activations a_0..a_5
constexpr uint32_t kDeltaOff = kActOff + 6 * 16384;   // 2 x 16 KB deltas (ping-pong)
constexpr uint32_t kBarOff = kDeltaOff + 2 * 16384;
constexpr uint32_t kGradSmemBytes = kBarOff + 256 + 1024;
constexpr uint32_t kColDW5 = 320, kColWork = 384;     // TMEM columns: dW_l at 64*l, dW_5^T at 320, working D at 384
constexpr int kGradThreads = 256;
} // namespace

__device__ __forceinline__ void store_row_sw128(uint8_t *tile, uint32_t row, const uint32_t o[32]) {
	uint8_t *r = tile + row * 128;
#pragma unroll
	for (int c = 0; c < 8; ++c)
		*(uint4 *)(r + ((c ^ (row & 7)) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
}
__device__ __forceinline__ void load_row_sw128(const uint8_t *tile, uint32_t row, uint32_t o[32]) {
	const uint8_t *r = tile + row * 128;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const uint4 t = *(const uint4 *)(r + ((c ^ (row & 7)) << 4));
		o[4 * c] = t.x, o[4 * c + 1] = t.y, o[4 * c + 2] = t.z, o[4 * c + 3] = t.w;
	}
}

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

template <int IN_MODE>
__global__ void __launch_bounds__(kGradThreads, 1)
    nrc_gradient_kernel(const GradParams p, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t *w_sm = smem + kWOff, *act_sm = smem + kActOff, *delta_sm = smem + kDeltaOff;
	uint64_t *bars = (uint64_t *)(smem + kBarOff);
	uint64_t *w_full = bars, *in_full = bars + 1, *d_full = bars + 2, *tile_done = bars + 3;
	uint32_t *tmem_slot = (uint32_t *)(bars + 4);
	float *red = (float *)(bars + 5); // 8 floats of block-reduction scratch

	// 8 warps: q = TMEM lane quarter (rows 32q..32q+31), h = which 32-column half of the 64-wide row this thread owns
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, h = warp >> 2;
	const uint32_t row = q * 32 + lane;
	uint64_t n = p.n;
	if (p.d_count) { // nrc_train_prepare.comp:17-18: count = min(count, NRC_TRAIN_BATCH_SIZE)
		const uint64_t c = *p.d_count;
		n = c < n ? c : n;
	}
	const uint32_t ntiles = (uint32_t)((n + NRC_TILE - 1) / NRC_TILE);
	float *my_partial = p.partials + (size_t)blockIdx.x * NRC_GRAD_STRIDE;
	if (blockIdx.x >= ntiles) { // nothing to do: contribute an all-zero partial so the reduction stays shape-stable
		for (uint32_t i = threadIdx.x; i < NRC_GRAD_STRIDE; i += blockDim.x)
			my_partial[i] = 0.0f;
		return;
	}
	const uint32_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

	if (threadIdx.x == 0) {
		mbar_init(w_full, 1), mbar_init(in_full, 1), mbar_init(d_full, 1), mbar_init(tile_done, 1);
		fence_mbar_init();
	}
	if (warp == 0)
		tmem_alloc(tmem_slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = *tmem_slot;

	constexpr uint32_t id_fwd64 = make_idesc_f16_f32(128, 64, false, false);
	constexpr uint32_t id_fwd16 = make_idesc_f16_f32(128, 16, false, false);
	constexpr uint32_t id_da = make_idesc_f16_f32(128, 64, false, true);  // A = delta K-major, B = W MN-major
	constexpr uint32_t id_dw64 = make_idesc_f16_f32(64, 64, true, true);  // A = delta MN-major, B = act MN-major
	constexpr uint32_t id_dw5t = make_idesc_f16_f32(64, 16, true, true);  // dW_5^T: A = a_5 MN-major, B = delta_5 MN-major
	const uint32_t w_a = smem_u32(w_sm), act_a = smem_u32(act_sm), del_a = smem_u32(delta_sm);
	const uint32_t d_issue = tmem + kColWork;                              // issuer's view of the working accumulator
	const uint32_t d_mine = tmem_addr(tmem, q * 32, kColWork + 32 * h);    // this thread's 32 columns of its row
	auto desc = [](uint32_t addr) { return make_smem_desc_sw128(addr, 0, 1024); };
	// The CTA-wide barrier hands a stored operand (generic-proxy smem writes already fenced to the async proxy by their
	// writers) and a drained accumulator to the issuing thread; there is no separate issue warp to wake up.
	auto cta_sync = [&]() {
		tc_fence_before();
		__syncthreads();
	};
	auto store_half_row = [&](uint8_t *tile, const uint32_t *o16) { // 16 packed pairs = 32 columns = 4 swizzled 16 B chunks
		uint8_t *r = tile + row * 128;
#pragma unroll
		for (int c = 0; c < 4; ++c)
			*(uint4 *)(r + (((4 * h + c) ^ (row & 7)) << 4)) = make_uint4(o16[4 * c], o16[4 * c + 1], o16[4 * c + 2], o16[4 * c + 3]);
	};

	if (warp == 0) {
		if (elect_one()) {
			tma_prefetch_desc(&tm_w);
			mbar_arrive_expect_tx(w_full, NRC_LAYERS * 8192);
			for (int l = 0; l < NRC_LAYERS; ++l)
				tma_load_2d(w_sm + l * 8192, &tm_w, 0, l * 64, w_full);
			if (IN_MODE == NRC_IN_ENCODED) {
				mbar_arrive_expect_tx(in_full, 16384);
				tma_load_2d(act_sm, &tm_in, 0, (int32_t)(blockIdx.x * NRC_TILE), in_full);
			}
		}
		__syncwarp();
	}

	uint32_t d_ph = 0;
	float loss_acc = 0.0f;
	uint32_t valid_rows = 0;
	for (uint32_t j = 0; j < my_tiles; ++j) {
		const uint32_t tile = blockIdx.x + j * gridDim.x;
		const uint64_t gi = (uint64_t)tile * NRC_TILE + row;
		const bool valid = gi < n;
		float tgt[3] = {0.0f, 0.0f, 0.0f};
		if (IN_MODE != NRC_IN_ENCODED) {
			if (j > 0) // the previous tile's dW_0 MMA still reads a_0
				mbar_wait(tile_done, (j - 1) & 1);
			if (h == 0) {
				uint32_t o[32];
#pragma unroll
				for (int i = 0; i < 32; ++i)
					o[i] = 0u; // nrc_gradient.comp:27-34: zero input + zero target => exactly zero contribution
				if (IN_MODE == NRC_IN_UNPACKED) {
					if (valid) {
						float in[14];
						const float2 *src = (const float2 *)((const uint8_t *)p.in + gi * p.in_stride_bytes);
#pragma unroll
						for (int i = 0; i < 7; ++i) {
							const float2 t = __ldg(src + i);
							in[2 * i] = t.x, in[2 * i + 1] = t.y;
						}
						encode_nrc(in, o);
					}
				} else { // NRC_IN_IMAGE_RANDOM (gradient.comp:47-49)
					uint32_t px = p.seed_x + (uint32_t)(gi % 128u), py = p.seed_y + (uint32_t)(gi / 128u);
					pcg2d(px, py);
					const float sc = 1.0f / (float)0xffffffffu;
					const float u = sc * (float)px, v = sc * (float)py;
					if (valid) {
						sample_bilinear_rgb(p.image_rgba8, p.image_w, p.image_h, u, v, tgt);
						encode_oneblob32(u, v, o);
					}
				}
				store_row_sw128(act_sm, row, o);
				fence_proxy_async_smem();
			}
			cta_sync();
		}
		if (h == 0 && valid && IN_MODE != NRC_IN_IMAGE_RANDOM) {
			if (p.target_is_f16) {
				const __half *t = (const __half *)((const uint8_t *)p.target + gi * p.target_stride_bytes);
				tgt[0] = __half2float(t[0]), tgt[1] = __half2float(t[1]), tgt[2] = __half2float(t[2]);
			} else {
				const float *t = (const float *)((const uint8_t *)p.target + gi * p.target_stride_bytes);
				tgt[0] = t[0], tgt[1] = t[1], tgt[2] = t[2];
			}
		}
		// ------------------------------------------------------------------------------------------ forward
#pragma unroll 1
		for (int l = 0; l < NRC_LAYERS; ++l) {
			if (warp == 0) {
				if (elect_one()) {
					if (j == 0 && l == 0)
						mbar_wait(w_full, 0);
					if (IN_MODE == NRC_IN_ENCODED && l == 0)
						mbar_wait(in_full, j & 1);
					tc_fence_after();
#pragma unroll
					for (int k = 0; k < 4; ++k)
						mma_ss(d_issue, desc(act_a + l * 16384 + k * 32), desc(w_a + l * 8192 + k * 32), l < 5 ? id_fwd64 : id_fwd16, k > 0);
					tc_commit(d_full);
				}
				__syncwarp();
			}
			mbar_wait(d_full, d_ph);
			d_ph ^= 1;
			tc_fence_after();
			if (l < NRC_HIDDEN_LAYERS) { // a_{l+1} = fp16(relu(D))
				uint32_t v[32], o[16];
				tmem_ld_x32(d_mine, v);
				tc_wait_ld();
#pragma unroll
				for (int i = 0; i < 16; ++i)
					o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
				store_half_row(act_sm + (l + 1) * 16384, o);
				fence_proxy_async_smem();
			} else if (h == 0) { // output layer + loss gradient (NN_nv.glsl:148-196)
				uint32_t yv[4];
				tmem_ld_x4(tmem_addr(tmem, q * 32, kColWork), yv);
				tc_wait_ld();
				float y[3], g[3]; // y is the fp16 network output widened to fp32, as NNOutput3 returns it
#pragma unroll
				for (int c = 0; c < 3; ++c)
					y[c] = __half2float(__float2half_rn(__uint_as_float(yv[c])));
				float den = 1.0f;
				if (p.loss_kind == NRC_LOSS_RELATIVE_L2_LUMINANCE) {
					const float lum = 0.299f * fmaxf(y[0], 0.0f) + 0.587f * fmaxf(y[1], 0.0f) + 0.114f * fmaxf(y[2], 0.0f);
					den = lum * lum + 0.01f;
				}
#pragma unroll
				for (int c = 0; c < 3; ++c) {
					const float d = y[c] - tgt[c];
					g[c] = p.loss_kind == NRC_LOSS_L2 ? 2.0f * d * p.loss_scale : 2.0f * p.loss_scale * d / den;
					if (valid)
						loss_acc += d * d / den;
				}
				if (!valid)
					g[0] = g[1] = g[2] = 0.0f;
				valid_rows += valid ? 1u : 0u;
				if (valid && p.y_out) {
					float *yo = (float *)p.y_out + 3 * gi;
					yo[0] = y[0], yo[1] = y[1], yo[2] = y[2];
				}
				uint8_t *r = delta_sm + row * 128; // delta_5: 16 fp16 = logical chunks 0 and 1 of the row
				*(uint4 *)(r + ((0 ^ (row & 7)) << 4)) = make_uint4(cvt_pack_f16x2(g[0], g[1]), cvt_pack_f16x2(g[2], 0.0f), 0u, 0u);
				*(uint4 *)(r + ((1 ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
				fence_proxy_async_smem();
			}
			cta_sync();
		}
		// ------------------------------------------------------------------------------------------ backward
		// per layer: dA first (critical path: delta_{l-1} = (delta_l W_l) * [a_l > 0]), then dW_l += delta_l^T a_l,
		// which the tensor pipe executes while the epilogue of dA runs.
#pragma unroll 1
		for (int l = 5; l >= 0; --l) {
			const uint32_t dl = del_a + ((5 - l) & 1) * 16384; // delta_l
			if (warp == 0) {
				if (elect_one()) {
					tc_fence_after();
					if (l == 5) {
						mma_ss(d_issue, desc(dl), desc(w_a + 5 * 8192), id_da, 0);
						tc_commit(d_full);
#pragma unroll
						for (int k = 0; k < 8; ++k)
							mma_ss(tmem + kColDW5, desc(act_a + 5 * 16384 + k * 2048), desc(dl + k * 2048), id_dw5t, (j > 0) || (k > 0));
					} else {
						if (l > 0) {
#pragma unroll
							for (int k = 0; k < 4; ++k)
								mma_ss(d_issue, desc(dl + k * 32), desc(w_a + l * 8192 + k * 2048), id_da, k > 0);
							tc_commit(d_full);
						}
#pragma unroll
						for (int k = 0; k < 8; ++k)
							mma_ss(tmem + 64 * l, desc(dl + k * 2048), desc(act_a + l * 16384 + k * 2048), id_dw64, (j > 0) || (k > 0));
						if (l == 0) {
							tc_commit(tile_done);
							if (IN_MODE == NRC_IN_ENCODED && j + 1 < my_tiles) { // a_0 is free once dW_0 has consumed it
								mbar_wait(tile_done, j & 1);
								mbar_arrive_expect_tx(in_full, 16384);
								tma_load_2d(act_sm, &tm_in, 0, (int32_t)((blockIdx.x + (j + 1) * gridDim.x) * NRC_TILE), in_full);
							}
						}
					}
				}
				__syncwarp();
			}
			if (l >= 1) { // delta_{l-1} = fp16(D) * [a_l > 0], NaN -> 0 (NN_nv.glsl:198-220, 240-242)
				mbar_wait(d_full, d_ph);
				d_ph ^= 1;
				tc_fence_after();
				uint32_t v[32], a[16], o[16];
				tmem_ld_x32(d_mine, v);
				{
					const uint8_t *r = act_sm + l * 16384 + row * 128;
#pragma unroll
					for (int c = 0; c < 4; ++c) {
						const uint4 t = *(const uint4 *)(r + (((4 * h + c) ^ (row & 7)) << 4));
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
				store_half_row(delta_sm + ((5 - (l - 1)) & 1) * 16384, o);
				fence_proxy_async_smem();
				cta_sync();
			}
		}
	}
	// ---- all tiles issued: drain the dW accumulators (M=64 TMEM layout: row r -> lane (r%16) + 32*(r/16))
	mbar_wait(tile_done, (my_tiles - 1) & 1);
	tc_fence_after();
#pragma unroll 1
	for (int l = 0; l < NRC_HIDDEN_LAYERS; ++l) {
		uint32_t v[32];
		tmem_ld_x32(tmem_addr(tmem, q * 32, 64 * l + 32 * h), v);
		tc_wait_ld();
		if (lane < 16) {
			float4 *dst = (float4 *)(my_partial + l * 4096 + (q * 16 + lane) * 64 + 32 * h);
#pragma unroll
			for (int i = 0; i < 8; ++i)
				dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
				                     __uint_as_float(v[4 * i + 3]));
		}
	}
	if (h == 0) {
		uint32_t v[4];
		tmem_ld_x4(tmem_addr(tmem, q * 32, kColDW5), v); // dW_5^T: lane <-> in, column <-> out
		tc_wait_ld();
		if (lane < 16)
#pragma unroll
			for (int o = 0; o < 3; ++o)
				my_partial[5 * 4096 + o * 64 + q * 16 + lane] = __uint_as_float(v[o]);
	}
	// loss / count slots: fixed-order block reduction (deterministic)
	float cnt = (float)valid_rows;
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) {
		loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
		cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
	}
	if (h == 0 && lane == 0)
		red[q] = loss_acc, red[4 + q] = cnt;
	tc_fence_before();
	__syncthreads();
	if (threadIdx.x == 0) {
		my_partial[NRC_GRAD_LOSS_SLOT] = (red[0] + red[1]) + (red[2] + red[3]);
		my_partial[NRC_GRAD_COUNT_SLOT] = (red[4] + red[5]) + (red[6] + red[7]);
	}
	for (uint32_t i = NRC_GRAD_COUNT_SLOT + 1 + threadIdx.x; i < NRC_GRAD_STRIDE; i += kGradThreads)
		my_partial[i] = 0.0f;
	if (warp == 0)
		tmem_dealloc(tmem, 512);
}

uint32_t gradient_max_partials(int sms) { return (uint32_t)sms; }

template <int IN_MODE>
static cudaError_t launch_grad_t(const GradParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, uint32_t grid, cudaStream_t stream) {
	auto kern = nrc_gradient_kernel<IN_MODE>;
	static bool configured = false;
	if (!configured) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGradSmemBytes);
		if (e != cudaSuccess)
			return e;
		configured = true;
	}
	kern<<<grid, kGradThreads, kGradSmemBytes, stream>>>(p, tm_w, tm_in);
	return cudaGetLastError();
}

cudaError_t launch_gradient(const GradParams &p, const CUtensorMap &tm_w, const CUtensorMap &tm_in, int sms, uint32_t *num_partials,
                            cudaStream_t stream) {
	const uint64_t ntiles = (p.n + NRC_TILE - 1) / NRC_TILE;
	uint32_t grid = (uint32_t)(ntiles < (uint64_t)sms ? ntiles : (uint64_t)sms);
	if (grid == 0)
		grid = 1; // still emits one all-zero partial
	*num_partials = grid;
	switch (p.in_mode) {
	case NRC_IN_ENCODED:
		return launch_grad_t<NRC_IN_ENCODED>(p, tm_w, tm_in, grid, stream);
	case NRC_IN_UNPACKED:
		return launch_grad_t<NRC_IN_UNPACKED>(p, tm_w, tm_in, grid, stream);
	case NRC_IN_IMAGE_RANDOM:
		return launch_grad_t<NRC_IN_IMAGE_RANDOM>(p, tm_w, tm_in, grid, stream);
	}
	return cudaErrorInvalidValue;
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
__device__ __forceinline__ void adam_update(const AdamParams &a, uint32_t i, float grad_sum, float count, const NrcOptimizerState &st) {
	float g = __fdiv_rn(__fdiv_rn(grad_sum, count), NRC_LOSS_SCALE);
	if (isnan(g) || isinf(g))
		g = 0.0f;
	NrcOptimizerEntry e = a.entries[i];
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
__device__ __forceinline__ void publish_state_if_last(const AdamParams &a, const NrcOptimizerState &st) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(a.done_counter, 1u) == gridDim.x - 1) {
			*a.opt_state = st;
			*a.done_counter = 0;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// Deterministic reduction of the per-CTA partials (+ optionally the fused optimizer step).
// Block = 64 elements x 4 partial groups: group g adds partials g, g+4, g+8, ... in that order with 8 loads in flight,
// then the four group sums are combined as (s0 + s1) + (s2 + s3). The order is fixed => bit-reproducible.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kReduceElems = 64, kReduceGroups = 4;
__global__ void __launch_bounds__(kReduceElems *kReduceGroups) reduce_partials_kernel(const ReduceParams p) {
	__shared__ float sm[kReduceGroups][kReduceElems];
	__shared__ float cnt_sm[8];
	const uint32_t e = threadIdx.x % kReduceElems, g = threadIdx.x / kReduceElems;
	const uint32_t i = blockIdx.x * kReduceElems + e;
	float count = 0.0f;
	if (p.fuse_adam) { // every block needs the batch's record count: integers < 2^24, so any summation order is exact
		float c = threadIdx.x < p.num_partials ? p.partials[(size_t)threadIdx.x * NRC_GRAD_STRIDE + NRC_GRAD_COUNT_SLOT] : 0.0f;
#pragma unroll
		for (int off = 16; off > 0; off >>= 1)
			c += __shfl_xor_sync(0xffffffffu, c, off);
		if ((threadIdx.x & 31) == 0)
			cnt_sm[threadIdx.x >> 5] = c;
	}
	float acc = 0.0f;
	if (i < p.limit) {
		const float *src = p.partials + i;
		uint32_t q = g;
		for (; q + 7 * kReduceGroups < p.num_partials; q += 8 * kReduceGroups) {
			float v[8];
#pragma unroll
			for (int u = 0; u < 8; ++u)
				v[u] = src[(size_t)(q + u * kReduceGroups) * NRC_GRAD_STRIDE];
#pragma unroll
			for (int u = 0; u < 8; ++u)
				acc += v[u];
		}
		for (; q < p.num_partials; q += kReduceGroups)
			acc += src[(size_t)q * NRC_GRAD_STRIDE];
	}
	sm[g][e] = acc;
	__syncthreads();
	if (p.fuse_adam)
		count = ((cnt_sm[0] + cnt_sm[1]) + (cnt_sm[2] + cnt_sm[3])) + ((cnt_sm[4] + cnt_sm[5]) + (cnt_sm[6] + cnt_sm[7]));
	NrcOptimizerState st{};
	const bool do_adam = p.fuse_adam && count > 0.0f; // nrc_optimize.comp:33-34 / nrc_train_prepare.comp:22
	if (do_adam)
		st = advance_state(*p.adam.opt_state);
	if (g == 0 && i < p.limit) {
		const float s = (sm[0][e] + sm[1][e]) + (sm[2][e] + sm[3][e]);
		p.gradients[i] = p.accumulate ? p.gradients[i] + s : s;
		if (do_adam && i < NRC_WEIGHT_COUNT)
			adam_update(p.adam, i, s, count, st);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0 && p.d_count) { // nrc_train_prepare.comp:17-19: write the clamped count back
		const uint32_t c = *p.d_count;
		*p.d_count = c < p.batch_cap ? c : p.batch_cap;
	}
	if (do_adam)
		publish_state_if_last(p.adam, st);
}

cudaError_t launch_reduce(const ReduceParams &p, cudaStream_t stream) {
	if (p.fuse_adam && p.num_partials > kReduceElems * kReduceGroups)
		return cudaErrorInvalidValue;
	reduce_partials_kernel<<<(NRC_GRAD_STRIDE + kReduceElems - 1) / kReduceElems, kReduceElems * kReduceGroups, 0, stream>>>(p);
	return cudaGetLastError();
}

// stand-alone optimizer step (used when an all-reduce sits between the reduction and the step)
__global__ void __launch_bounds__(128) adam_kernel(const AdamParams a) {
	const float count = a.gradients[NRC_GRAD_COUNT_SLOT];
	if (!(count > 0.0f)) // nrc_optimize.comp:33-34 / nrc_train_prepare.comp:22
		return;
	const NrcOptimizerState st = advance_state(*a.opt_state);
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < NRC_WEIGHT_COUNT)
		adam_update(a, i, a.gradients[i], count, st);
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
