// probe_mma_rate.cu -- tcgen05.mma issue-rate / dependent-chain latency microbenchmark (development tool).
// One CTA per SM; warp 1's elected thread issues `iters` groups of `chain` MMAs (K=16 each). Each group accumulates
// into accumulator (g % nacc); so nacc = 1 is a fully dependent stream, nacc >= 2 interleaves independent chains when
// interleave=1 (k-step-major order across nacc accumulators).
#include "sm100_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace sm100;

template <int N, bool TS, bool INTERLEAVE>
__global__ void __launch_bounds__(128, 1) k(int iters, int nacc, long long *out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const uint32_t warp = threadIdx.x >> 5;
	for (uint32_t i = threadIdx.x * 4; i < 65536; i += 512)
		*(uint32_t *)(smem + i) = 0x3c003c00u;
	fence_proxy_async_smem();
	if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	if (warp == 0) tmem_alloc(&slot, 512);
	tc_fence_before(); __syncthreads(); tc_fence_after();
	const uint32_t tb = slot;
	if (warp == 1) {
		if (elect_one()) {
			constexpr uint32_t idesc = make_idesc_f16_f32(128, N, false, false);
			const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
			const uint32_t a_t = tb + 480; // 32 columns of A operand
			long long t0 = clock64();
			if (!INTERLEAVE) {
				for (int it = 0; it < iters; ++it) {
					const uint32_t d = tb + (it % nacc) * N;
#pragma unroll
					for (int kk = 0; kk < 4; ++kk) {
						if (TS) mma_ts(d, a_t + kk * 8, make_smem_desc_sw128(sb + kk * 32, 0, 1024), idesc, kk > 0);
						else mma_ss(d, make_smem_desc_sw128(sa + kk * 32, 0, 1024), make_smem_desc_sw128(sb + kk * 32, 0, 1024), idesc, kk > 0);
					}
				}
			} else {
				for (int it = 0; it < iters; it += nacc) {
#pragma unroll
					for (int kk = 0; kk < 4; ++kk)
						for (int a = 0; a < nacc; ++a) {
							const uint32_t d = tb + a * N;
							if (TS) mma_ts(d, a_t + kk * 8, make_smem_desc_sw128(sb + kk * 32, 0, 1024), idesc, kk > 0);
							else mma_ss(d, make_smem_desc_sw128(sa + kk * 32, 0, 1024), make_smem_desc_sw128(sb + kk * 32, 0, 1024), idesc, kk > 0);
						}
				}
			}
			tc_commit(&bar);
			mbar_wait(&bar, 0);
			long long t1 = clock64();
			if (blockIdx.x == 0) out[0] = t1 - t0;
		}
		__syncwarp();
	}
	tc_fence_before(); __syncthreads();
	if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, bool TS, bool IL> void run(const char *name, int nacc, long long *dc) {
	const int iters = 4800;
	cudaFuncSetAttribute(k<N, TS, IL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
	k<N, TS, IL><<<148, 128, 66560>>>(iters, nacc, dc);
	cudaError_t e = cudaDeviceSynchronize();
	long long c = 0;
	cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
	printf("%-34s nacc=%d : %7.1f cycles per MMA instr (%s)\n", name, nacc, (double)c / (iters * 4.0), cudaGetErrorString(e));
}

int main() {
	long long *dc;
	cudaMalloc(&dc, 8);
	run<64, true, false>("TS N=64 chain4 sequential", 1, dc);
	run<64, true, false>("TS N=64 chain4 sequential", 2, dc);
	run<64, true, false>("TS N=64 chain4 sequential", 4, dc);
	run<64, true, true>("TS N=64 k-major interleaved", 2, dc);
	run<64, true, true>("TS N=64 k-major interleaved", 3, dc);
	run<64, true, true>("TS N=64 k-major interleaved", 4, dc);
	run<64, true, true>("TS N=64 k-major interleaved", 6, dc);
	run<64, false, false>("SS N=64 chain4 sequential", 1, dc);
	run<64, false, true>("SS N=64 k-major interleaved", 4, dc);
	run<16, true, false>("TS N=16 chain4 sequential", 1, dc);
	run<16, true, true>("TS N=16 k-major interleaved", 4, dc);
	run<32, true, false>("TS N=32 chain4 sequential", 1, dc);
	run<128, true, false>("TS N=128 chain4 sequential", 1, dc);
	run<128, true, true>("TS N=128 k-major interleaved", 3, dc);
	run<256, true, false>("TS N=256 chain4 sequential", 1, dc);
	run<256, false, false>("SS N=256 chain4 sequential", 1, dc);
	return 0;
}
