// probe_dw_rate.cu -- development tool: tensor-pipe cost of one step of the training kernel's two-tile pipeline (forward layer:
// 4 x TS M=128 N=64, dA: 4 x TS with W read MN-major, dW: 8 x SS M=64 N=64 with both operands MN-major) in isolation, and what
// the epilogue warps' shared-memory traffic (16-byte swizzled row stores + fence.proxy.async, optional mask loads) costs while
// those MMAs run. One CTA per SM; warp 8 issues, warps 0..7 store.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ivknrc_b200/csrc tools/probe_dw_rate.cu -o tools/probe_dw_rate
#include "sm100_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace sm100;

// what: bit 0 forward, bit 1 dA, bit 2 dW ; stores: 0 none, 1 = one 16 KB tile store + fence per iteration, 2 = two stores + one
// 16 KB load per iteration (what a step's epilogues do)
__global__ void __launch_bounds__(288, 1) k(int steps, int what, int stores, int store_iters, long long *out) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	__shared__ uint64_t bar, bar2;
	__shared__ uint32_t slot;
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (uint32_t i = threadIdx.x * 4; i < 196608; i += 288 * 4)
		*(uint32_t *)(smem + i) = 0x3c003c00u;
	fence_proxy_async_smem();
	if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 8); fence_mbar_init(); }
	if (warp == 8) tmem_alloc(&slot, 512);
	tc_fence_before(); __syncthreads(); tc_fence_after();
	const uint32_t tb = slot;
	uint8_t *w_sm = smem, *pool = smem + 49152; // weights 48 KB, then 16 KB tiles
	if (warp == 8) {
		if (elect_one()) {
			constexpr uint32_t id_fwd64 = make_idesc_f16_f32(128, 64, false, false);
			constexpr uint32_t id_da = make_idesc_f16_f32(128, 64, false, true);
			constexpr uint32_t id_dw64 = make_idesc_f16_f32(64, 64, true, true);
			const uint32_t w_desc = smem_desc_lo(smem_u32(w_sm)), pool_desc = smem_desc_lo(smem_u32(pool));
			constexpr uint32_t dhi = kSmemDescHiSw128;
			long long t0 = clock64();
			for (int s = 0; s < steps; ++s) {
				const uint32_t l = (uint32_t)(s % 5);
				if (what & 1)
#pragma unroll
					for (int kk = 0; kk < 4; ++kk)
						mma_ts_lh(tb + 192, tb + 320 + kk * 8, w_desc + l * 512 + kk * 2, dhi, id_fwd64, kk > 0);
				if (what & 2)
#pragma unroll
					for (int kk = 0; kk < 4; ++kk)
						mma_ts_lh(tb + 256, tb + 352 + kk * 8, w_desc + l * 512 + kk * 128, dhi, id_da, kk > 0);
				if (what & 4)
#pragma unroll
					for (int kk = 0; kk < 8; ++kk)
						mma_ss_lh(tb + 64 * (l >> 1) + ((16 * (l & 1)) << 16), pool_desc + 7 * 1024 + kk * 128, pool_desc + l * 1024 + kk * 128, dhi, id_dw64, 1);
			}
			tc_commit(&bar);
			mbar_wait(&bar, 0);
			long long t1 = clock64();
			if (blockIdx.x == 0) out[0] = t1 - t0;
		}
		__syncwarp();
	} else if (stores) {
		const uint32_t q = warp & 3, h = warp >> 2, row = q * 32 + lane;
		uint32_t acc = 0;
		asm volatile("bar.sync 1, 256;");
		long long t0 = clock64();
		for (int it = 0; it < store_iters; ++it) {
			uint8_t *r = pool + 8 * 16384 + (it & 1) * 16384 + row * 128;
			if (stores >= 3) { // 3: stores only, 4: fence only, 5: stores + one mbarrier arrival per warp (no proxy fence)
				if (stores != 4)
#pragma unroll
					for (int c = 0; c < 4; ++c)
						asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(r + (((4 * h + c) ^ (row & 7)) << 4))), "r"(it), "r"(acc), "r"(c), "r"(it) : "memory");
				if (stores == 4)
					fence_proxy_async_smem();
				if (stores == 5) {
					__syncwarp();
					if (lane == 0)
						mbar_arrive(&bar2);
				}
				__syncwarp();
				continue;
			}
			if (stores >= 2) {
				const uint8_t *rr = pool + (it % 5) * 16384 + row * 128;
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					uint4 t;
					asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(smem_u32(rr + (((4 * h + c) ^ (row & 7)) << 4))));
					acc += t.x ^ t.y ^ t.z ^ t.w;
				}
			}
			for (int rep = 0; rep < (stores >= 2 ? 2 : 1); ++rep) {
#pragma unroll
				for (int c = 0; c < 4; ++c)
					asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(r + (((4 * h + c) ^ (row & 7)) << 4))), "r"(it), "r"(acc), "r"(c), "r"(rep) : "memory");
				fence_proxy_async_smem();
			}
			__syncwarp();
		}
		long long t1 = clock64();
		if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = t1 - t0 + (acc == 0xdeadbeefu);
	}
	tc_fence_before(); __syncthreads();
	if (warp == 8) tmem_dealloc(tb, 512);
}

int main() {
	long long *dc;
	cudaMalloc(&dc, 16);
	cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000);
	auto run = [&](const char *name, int what, int stores, int steps, int store_iters) {
		cudaMemset(dc, 0, 16);
		k<<<148, 288, 230000>>>(steps, what, stores, store_iters, dc);
		cudaError_t e = cudaDeviceSynchronize();
		long long c[2] = {0, 0};
		cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
		printf("%-46s: %7.1f cycles per step (issuer)", name, steps ? (double)c[0] / steps : 0.0);
		if (stores) printf(" | %7.1f cycles per epilogue-store iteration", (double)c[1] / store_iters);
		printf(" (%s)\n", cudaGetErrorString(e));
	};
	const int S = 4000;
	run("forward only (4 TS)", 1, 0, S, 0);
	run("dA only (4 TS, B MN-major)", 2, 0, S, 0);
	run("dW only (8 SS M=64, MN/MN)", 4, 0, S, 0);
	run("forward + dA", 3, 0, S, 0);
	run("forward + dA + dW (one step)", 7, 0, S, 0);
	run("no MMA, 1 tile store + fence / iter", 0, 1, 0, S);
	run("no MMA, 2 stores + 1 load / iter", 0, 2, 0, S);
	run("no MMA, 1 tile store, no fence", 0, 3, 0, S);
	run("no MMA, fence.proxy.async only", 0, 4, 0, S);
	run("no MMA, 1 tile store + mbarrier arrive", 0, 5, 0, S);
	run("fwd + dA + dW  with 1 store, no fence", 7, 3, 3 * S, S);
	run("fwd + dA + dW  with 1 store + mbarrier arrive", 7, 5, 3 * S, S);
	// stores running under a long MMA stream (steps chosen so that the MMAs outlast the stores)
	run("fwd + dA + dW  with 1 store / iter", 7, 1, 3 * S, S);
	run("fwd + dA + dW  with 2 stores + load / iter", 7, 2, 3 * S, S);
	run("dW only        with 2 stores + load / iter", 4, 2, 6 * S, S);
	run("fwd + dA only  with 2 stores + load / iter", 3, 2, 8 * S, S);
	// ... and the MMA stream's pace while the stores outlast it
	run("fwd + dA + dW  UNDER 1 store / iter", 7, 1, S, 8 * S);
	run("fwd + dA + dW  UNDER 2 stores + load / iter", 7, 2, S, 8 * S);
	run("dW only        UNDER 2 stores + load / iter", 4, 2, S, 6 * S);
	return 0;
}
