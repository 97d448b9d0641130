// probe_tmem.cu -- development tool: does tcgen05.ld / tcgen05.st traffic of epilogue warps slow tcgen05.mma down
// (and vice versa)? One CTA per SM; warp 0 issues dependent chains of M=128 N=64 K=16 TS-form MMAs round-robin over 4
// accumulators; EW epilogue warps (4 per TMEM lane quarter) run LDTM / STTM loops on other columns.
#include "../vknrc_b200/csrc/sm100_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace sm100;

__device__ __forceinline__ void tmem_ld_x32_pack16(uint32_t taddr, uint32_t *v) { // 32 registers <- 64 columns of 16-bit data
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	             : SM100_R16(v, 0), SM100_R16(v, 16)
	             : "r"(taddr)
	             : "memory");
}
__device__ __forceinline__ void tmem_ld_x16_pack16(uint32_t taddr, uint32_t *v) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	             : SM100_R16(v, 0)
	             : "r"(taddr)
	             : "memory");
}

// LD_KIND: 0 none, 1 = 2 x LDTM.x32 (64 fp32 columns) per iteration, 2 = LDTM.x32.pack16 (64 columns of 16-bit -> 32 regs),
//          4 = 2 x LDTM.x32 + STTM.x32 (the real epilogue's traffic), 5 = STTM.x32 only
template <int LD_KIND, bool F16ACC>
__global__ void __launch_bounds__(32 * 21, 1) k(int mma_iters, int ld_iters, int ew, long long *out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const uint32_t warp = threadIdx.x >> 5;
	for (uint32_t i = threadIdx.x * 4; i < 65536; i += blockDim.x * 4)
		*(uint32_t *)(smem + i) = 0x3c003c00u;
	fence_proxy_async_smem();
	if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
	if (warp == 0) tmem_alloc(&slot, 512);
	tc_fence_before(); __syncthreads(); tc_fence_after();
	const uint32_t tb = slot;
	if (warp == 0) {
		if (mma_iters > 0 && elect_one()) {
			constexpr uint32_t idesc = F16ACC ? (make_idesc_f16_f32(128, 64, false, false) & ~(3u << 4)) : make_idesc_f16_f32(128, 64, false, false);
			const uint32_t sb = smem_u32(smem + 32768);
			const uint32_t a_t = tb + 480;
			long long t0 = clock64();
			for (int it = 0; it < mma_iters; ++it) {
				const uint32_t d = tb + (it & 3) * 64;
#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
					mma_ts(d, a_t + kk * 8, make_smem_desc_sw128(sb + kk * 32, 0, 1024), idesc, kk > 0);
			}
			tc_commit(&bar);
			mbar_wait(&bar, 0);
			long long t1 = clock64();
			if (blockIdx.x == 0) out[0] = t1 - t0;
		}
		__syncwarp();
	} else if (warp <= (uint32_t)ew && LD_KIND != 0) {
		const uint32_t q = warp & 3;
		const uint32_t base = tmem_addr(tb, q * 32, 256); // columns 256.. (not touched by the MMAs)
		uint32_t v[64];
#pragma unroll
		for (int i = 0; i < 64; ++i) v[i] = i;
		uint32_t acc = 0;
		long long t0 = clock64();
#pragma unroll 1
		for (int it = 0; it < ld_iters; ++it) {
			const uint32_t c = base + (it & 1) * 96;
			if (LD_KIND == 1 || LD_KIND == 4) {
				tmem_ld_x32(c, v);
				tmem_ld_x32(c + 32, v + 32);
				tc_wait_ld();
			} else if (LD_KIND == 2) {
				tmem_ld_x32_pack16(c, v);
				tc_wait_ld();
			}
			if (LD_KIND == 4 || LD_KIND == 5) {
#pragma unroll
				for (int i = 0; i < 32; ++i) v[i] ^= v[32 + i];
				tmem_st_x32(c + 64, v);
				tc_wait_st();
			}
			acc += v[0] + v[63];
		}
		long long t1 = clock64();
		if (blockIdx.x == 0 && threadIdx.x == 32) out[1] = t1 - t0;
		if (acc == 0x12345678u) out[2] = acc;
	}
	tc_fence_before(); __syncthreads();
	if (warp == 0) tmem_dealloc(tb, 512);
}

template <int LD_KIND, bool F16ACC> void run(const char *name, int mma_iters, int ld_iters, int ew, long long *dc) {
	cudaFuncSetAttribute(k<LD_KIND, F16ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
	cudaMemset(dc, 0, 32);
	k<LD_KIND, F16ACC><<<148, 32 * 21, 66560>>>(mma_iters, ld_iters, ew, dc);
	cudaError_t e = cudaDeviceSynchronize();
	long long c[2] = {0, 0};
	cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
	printf("%-44s ew=%2d : %6.1f cyc/MMA   %7.1f cyc/epilogue-iter/warp  (%s)\n", name, ew, mma_iters ? (double)c[0] / (mma_iters * 4.0) : 0.0,
	       ld_iters ? (double)c[1] / ld_iters : 0.0, cudaGetErrorString(e));
}

int main() {
	long long *dc;
	cudaMalloc(&dc, 32);
	const int M = 4000, L = 4000;
	run<0, false>("MMA only (fp32 acc)", M, 0, 0, dc);
	run<0, true>("MMA only (fp16 acc)", M, 0, 0, dc);
	for (int ew : {4, 8, 20}) {
		run<1, false>("LD 2x x32 only", 0, L, ew, dc);
		run<1, false>("LD 2x x32 + MMA", M * 4, L, ew, dc);
		run<2, false>("LD x32.pack16 (64 cols) only", 0, L, ew, dc);
		run<2, true>("LD x32.pack16 + MMA(f16 acc)", M * 4, L, ew, dc);
		run<5, false>("ST x32 only", 0, L, ew, dc);
		run<4, false>("LD 2x x32 + ST x32 only", 0, L, ew, dc);
		run<4, false>("LD 2x x32 + ST x32 + MMA", M * 4, L, ew, dc);
	}
	return 0;
}
