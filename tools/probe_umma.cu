// probe_umma.cu -- stand-alone sm_100a probe (development tool, not part of the product or the tests).
//
//  (1) checks every tcgen05.mma operand form the NRC kernels rely on against a CPU GEMM:
//        SS K-major/K-major, TS (A in TMEM), B MN-major, A+B MN-major with M=64, N=16 variants;
//  (2) measures what bounds the fused MLP: MMA issue rate (SS vs TS), tcgen05.ld/st rate, cvt.relu rate and the
//      combined layer epilogue with 4/8/16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I vknrc_b200/csrc tools/probe_umma.cu -o tools/probe_umma
#include "sm100_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <string>

using namespace sm100;

#define CK(x)                                                                                                          \
	do {                                                                                                               \
		cudaError_t e_ = (x);                                                                                          \
		if (e_ != cudaSuccess) {                                                                                       \
			printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                           \
			exit(2);                                                                                                   \
		}                                                                                                              \
	} while (0)

struct ProbeParams {
	const uint8_t *a_img;
	const uint8_t *b_img;
	const uint32_t *a_rows; // [128][32] packed pairs, used when a_tmem
	uint32_t a_bytes, b_bytes;
	int a_tmem;
	uint32_t idesc;
	uint32_t a_lbo, a_sbo, a_kstep;
	uint32_t b_lbo, b_sbo, b_kstep;
	int ksteps, ncols_out;
	float *d_out; // [128][ncols_out]
};

__global__ void __launch_bounds__(128, 1) probe_gemm(ProbeParams p) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t tmem_base_slot;
	uint8_t *sa = smem, *sb = smem + 32768;
	const uint32_t tid = threadIdx.x, warp = tid >> 5;
	for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16)
		*(uint4 *)(sa + i) = *(const uint4 *)(p.a_img + i);
	for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16)
		*(uint4 *)(sb + i) = *(const uint4 *)(p.b_img + i);
	fence_proxy_async_smem();
	if (tid == 0) {
		mbar_init(&bar, 1);
		fence_mbar_init();
	}
	if (warp == 0)
		tmem_alloc(&tmem_base_slot, 256);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tbase = tmem_base_slot;
	const uint32_t a_t = tbase + 128; // A operand columns [128, 160)
	if (p.a_tmem) {
		uint32_t v[32];
		for (int i = 0; i < 32; ++i)
			v[i] = p.a_rows[tid * 32 + i];
		tmem_st_x32(tmem_addr(a_t, warp * 32, 0), v);
		tc_wait_st();
	}
	// zero-fill D so that untouched lanes/columns are recognisable
	{
		uint32_t z[32];
		for (int i = 0; i < 32; ++i)
			z[i] = 0x7fc00000u; // NaN pattern
		tmem_st_x32(tmem_addr(tbase, warp * 32, 0), z);
		tmem_st_x32(tmem_addr(tbase, warp * 32, 32), z);
		tc_wait_st();
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	if (tid == 0) {
		for (int k = 0; k < p.ksteps; ++k) {
			uint64_t bd = make_smem_desc_sw128(smem_u32(sb) + k * p.b_kstep, p.b_lbo, p.b_sbo);
			if (p.a_tmem)
				mma_ts(tbase, a_t + k * p.a_kstep, bd, p.idesc, k > 0);
			else
				mma_ss(tbase, make_smem_desc_sw128(smem_u32(sa) + k * p.a_kstep, p.a_lbo, p.a_sbo), bd, p.idesc, k > 0);
		}
		tc_commit(&bar);
	}
	mbar_wait(&bar, 0);
	tc_fence_after();
	for (int c = 0; c < p.ncols_out; c += 16) {
		uint32_t v[16];
		tmem_ld_x16(tmem_addr(tbase, warp * 32, c), v);
		tc_wait_ld();
		for (int i = 0; i < 16; ++i)
			p.d_out[(size_t)tid * p.ncols_out + c + i] = __uint_as_float(v[i]);
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0)
		tmem_dealloc(tbase, 256);
}

// ------------------------------------------------------------------------------------------------ host side of (1)
static uint16_t f2h(float f) {
	__half h = __float2half_rn(f);
	uint16_t u;
	memcpy(&u, &h, 2);
	return u;
}
static float h2f(uint16_t u) {
	__half h;
	memcpy(&h, &u, 2);
	return __half2float(h);
}

struct Case {
	std::string name;
	int M, N, K;
	bool a_tmem, a_mn, b_mn;
	uint32_t a_lbo = 0, b_lbo = 0; // bytes
};

// tile image: rows of 128 B, swizzled. K-major: row = mn index, col = k. MN-major: row = k, col = mn index.
static void put(std::vector<uint8_t> &img, int row, int col, uint16_t v) {
	uint32_t off = sw128_offset(row, col);
	if (off + 2 > img.size())
		img.resize(off + 2, 0);
	memcpy(&img[off], &v, 2);
}

static bool run_case(const Case &c) {
	std::vector<uint16_t> A(c.M * c.K), B(c.N * c.K);
	srand(1234 + c.M * 7 + c.N * 3 + c.K);
	for (auto &v : A)
		v = f2h((float)(rand() % 17 - 8) / 8.0f);
	for (auto &v : B)
		v = f2h((float)(rand() % 13 - 6) / 4.0f);
	std::vector<uint8_t> aimg(32768, 0), bimg(16384, 0);
	std::vector<uint32_t> arows(128 * 32, 0);
	for (int m = 0; m < c.M; ++m)
		for (int k = 0; k < c.K; ++k) {
			if (c.a_tmem) {
				uint32_t &w = arows[m * 32 + k / 2];
				w |= (uint32_t)A[m * c.K + k] << ((k & 1) * 16);
			} else if (c.a_mn)
				put(aimg, k, m, A[m * c.K + k]);
			else
				put(aimg, m, k, A[m * c.K + k]);
		}
	for (int n = 0; n < c.N; ++n)
		for (int k = 0; k < c.K; ++k) {
			if (c.b_mn)
				put(bimg, k, n, B[n * c.K + k]);
			else
				put(bimg, n, k, B[n * c.K + k]);
		}
	aimg.resize(32768, 0);
	bimg.resize(16384, 0);
	ProbeParams p{};
	uint8_t *da, *db;
	uint32_t *dr;
	float *dd;
	int ncols = (c.N + 15) / 16 * 16;
	CK(cudaMalloc(&da, 32768));
	CK(cudaMalloc(&db, 16384));
	CK(cudaMalloc(&dr, 128 * 32 * 4));
	CK(cudaMalloc(&dd, 128 * ncols * 4));
	CK(cudaMemcpy(da, aimg.data(), 32768, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(db, bimg.data(), 16384, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dr, arows.data(), 128 * 32 * 4, cudaMemcpyHostToDevice));
	p.a_img = da, p.b_img = db, p.a_rows = dr, p.a_bytes = 32768, p.b_bytes = 16384;
	p.a_tmem = c.a_tmem;
	p.idesc = make_idesc_f16_f32(c.M, c.N, c.a_mn, c.b_mn);
	p.a_lbo = c.a_lbo, p.a_sbo = 1024, p.a_kstep = c.a_tmem ? 8 : (c.a_mn ? 2048 : 32);
	p.b_lbo = c.b_lbo, p.b_sbo = 1024, p.b_kstep = c.b_mn ? 2048 : 32;
	p.ksteps = c.K / 16, p.ncols_out = ncols, p.d_out = dd;
	CK(cudaFuncSetAttribute(probe_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
	probe_gemm<<<1, 128, 49152>>>(p);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		printf("CASE %-28s : CUDA ERROR %s\n", c.name.c_str(), cudaGetErrorString(e));
		exit(3); // context is dead after a trap
	}
	std::vector<float> D(128 * ncols);
	CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
	double maxerr = 0;
	int bad = 0;
	for (int m = 0; m < c.M; ++m) {
		int lane = (c.M == 128) ? m : (m % 16) + 32 * (m / 16);
		for (int n = 0; n < c.N; ++n) {
			double ref = 0;
			for (int k = 0; k < c.K; ++k)
				ref += (double)h2f(A[m * c.K + k]) * (double)h2f(B[n * c.K + k]);
			double got = D[lane * ncols + n];
			double err = fabs(got - ref);
			if (!(err <= 1e-3 * (1 + fabs(ref))))
				++bad;
			if (err > maxerr || err != err)
				maxerr = err;
		}
	}
	printf("CASE %-28s M=%3d N=%3d K=%3d : %s  (bad=%d maxerr=%g)  D[0][0..3]= %g %g %g %g\n", c.name.c_str(), c.M, c.N,
	       c.K, bad ? "FAIL" : "ok", bad, maxerr, D[0], D[1], D[2], D[3]);
	cudaFree(da), cudaFree(db), cudaFree(dr), cudaFree(dd);
	return bad == 0;
}

// ------------------------------------------------------------------------------------------------ (2) micro-benchmarks
// mode 0: SS M128 N64 K-major/K-major; 1: TS M128 N64; 2: SS M64 N64 MN/MN (dW form); 3: SS M128 N64 A K-major B MN-major
// 4: TS M128 N16; 5: SS M128 N128 (for reference)
__global__ void __launch_bounds__(128, 1) bench_mma(int mode, int iters, long long *cycles_out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint64_t bar;
	__shared__ uint32_t slot;
	const uint32_t tid = threadIdx.x, warp = tid >> 5;
	for (uint32_t i = tid * 4; i < 49152; i += 512)
		*(uint32_t *)(smem + i) = 0x3c003c00u; // fp16 1.0
	fence_proxy_async_smem();
	if (tid == 0) {
		mbar_init(&bar, 1);
		fence_mbar_init();
	}
	if (warp == 0)
		tmem_alloc(&slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tb = slot;
	if (tid == 0) {
		uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
		uint32_t idesc;
		int ksteps = 4;
		if (mode == 0 || mode == 1) idesc = make_idesc_f16_f32(128, 64, false, false);
		else if (mode == 2) idesc = make_idesc_f16_f32(64, 64, true, true), ksteps = 8;
		else if (mode == 3) idesc = make_idesc_f16_f32(128, 64, false, true);
		else if (mode == 4) idesc = make_idesc_f16_f32(128, 16, false, false);
		else idesc = make_idesc_f16_f32(128, 128, false, false);
		long long t0 = clock64();
		for (int it = 0; it < iters; ++it) {
			uint32_t d = tb + (it & 1) * 128;
			for (int k = 0; k < ksteps; ++k) {
				if (mode == 1 || mode == 4)
					mma_ts(d, tb + 256 + k * 8, make_smem_desc_sw128(sb + k * 32, 0, 1024), idesc, k > 0);
				else if (mode == 2)
					mma_ss(d, make_smem_desc_sw128(sa + k * 2048, 0, 1024), make_smem_desc_sw128(sb + (k & 3) * 2048, 0, 1024), idesc, k > 0);
				else if (mode == 3)
					mma_ss(d, make_smem_desc_sw128(sa + k * 32, 0, 1024), make_smem_desc_sw128(sb + k * 2048, 0, 1024), idesc, k > 0);
				else
					mma_ss(d, make_smem_desc_sw128(sa + k * 32, 0, 1024), make_smem_desc_sw128(sb + k * 32, 0, 1024), idesc, k > 0);
			}
		}
		tc_commit(&bar);
		mbar_wait(&bar, 0);
		long long t1 = clock64();
		if (blockIdx.x == 0)
			cycles_out[0] = t1 - t0;
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 0)
		tmem_dealloc(tb, 512);
}

// epilogue micro-benchmarks. mode 0: ld x64 only; 1: ld x32 twice; 2: cvt.relu only (64 -> 32); 3: st x32 only;
// 4: ld x64 + cvt + st x32 (the layer epilogue); 5: same + sts to smem (8 x 16 B, swizzled) instead of tcgen05.st
__global__ void __launch_bounds__(512, 1) bench_epi(int mode, int iters, long long *cycles_out, uint32_t *sink) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint32_t slot;
	const uint32_t tid = threadIdx.x, warp = tid >> 5;
	if (warp == 0)
		tmem_alloc(&slot, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tb = slot;
	const uint32_t lane_base = (warp & 3) * 32, grp = warp >> 2; // each group of 4 warps owns 96 columns
	const uint32_t dcol = grp * 96, acol = grp * 96 + 64;
	uint32_t v[64], o[32];
	for (int i = 0; i < 64; ++i)
		v[i] = tid * 64 + i;
	for (int i = 0; i < 32; ++i)
		o[i] = i;
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		if (mode == 0 || mode == 4 || mode == 5) {
			tmem_ld_x64(tmem_addr(tb, lane_base, dcol), v);
			tc_wait_ld();
		}
		if (mode == 1) {
			tmem_ld_x32(tmem_addr(tb, lane_base, dcol), v);
			tmem_ld_x32(tmem_addr(tb, lane_base, dcol + 32), v + 32);
			tc_wait_ld();
		}
		if (mode == 2 || mode == 4 || mode == 5) {
#pragma unroll
			for (int i = 0; i < 32; ++i)
				o[i] = cvt_relu_pack_f16x2(__uint_as_float(v[2 * i] + (mode == 2 ? o[i] : 0)), __uint_as_float(v[2 * i + 1]));
		}
		if (mode == 3 || mode == 4) {
			tmem_st_x32(tmem_addr(tb, lane_base, acol), o);
			tc_wait_st();
		}
		if (mode == 5) {
			uint8_t *row = smem + grp * 16384 + (tid & 127) * 128;
#pragma unroll
			for (int c = 0; c < 8; ++c)
				*(uint4 *)(row + ((c ^ (tid & 7)) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
			fence_proxy_async_smem();
		}
	}
	long long t1 = clock64();
	uint32_t acc = 0;
	for (int i = 0; i < 32; ++i)
		acc += o[i] + v[i] + v[i + 32];
	sink[blockIdx.x * blockDim.x + tid] = acc;
	if (blockIdx.x == 0 && tid == 0)
		cycles_out[0] = t1 - t0;
	tc_fence_before();
	__syncthreads();
	if (warp == 0)
		tmem_dealloc(tb, 512);
}

int main(int argc, char **argv) {
	int dev = 0;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, dev));
	printf("device: %s sm_%d%d SMs=%d clock=%d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
	bool only_bench = argc > 1 && !strcmp(argv[1], "bench");
	if (!only_bench) {
		std::vector<Case> cases = {
		    {"fwd  SS Kmaj x Kmaj", 128, 64, 64, false, false, false},
		    {"fwd  TS tmemA x Kmaj", 128, 64, 64, true, false, false},
		    {"out  SS N=16", 128, 16, 64, false, false, false},
		    {"out  TS N=16", 128, 16, 64, true, false, false},
		    {"dA   SS Kmaj x MNmaj", 128, 64, 64, false, false, true},
		    {"dA   TS tmemA x MNmaj", 128, 64, 64, true, false, true},
		    {"dA5  SS K=16 x MNmaj", 128, 64, 16, false, false, true},
		    {"dA5  TS K=16 x MNmaj", 128, 64, 16, true, false, true},
		    {"dW   SS M64 MN x MN K128", 64, 64, 128, false, true, true},
		    {"dW5t SS M64 MN x MN N16", 64, 16, 128, false, true, true},
		    {"dW   SS M64 MN x MN lbo", 64, 64, 128, false, true, true, 8192, 8192},
		    {"fwd  SS M64", 64, 64, 64, false, false, false},
		};
		int fails = 0;
		if (argc > 2 && !strcmp(argv[1], "case")) { // one case per process: a trap kills only that case
			int i = atoi(argv[2]);
			if (i < 0 || i >= (int)cases.size())
				return 9;
			return run_case(cases[i]) ? 0 : 1;
		}
		for (auto &c : cases)
			fails += !run_case(c);
		printf("probe cases failed: %d\n", fails);
	}
	// ---- benchmarks
	long long *dc;
	uint32_t *sink;
	CK(cudaMalloc(&dc, 8));
	CK(cudaMalloc(&sink, 148 * 512 * 4));
	CK(cudaFuncSetAttribute(bench_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
	CK(cudaFuncSetAttribute(bench_epi, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
	const char *mnames[] = {"SS M128 N64 (4 k-steps)", "TS M128 N64 (4 k-steps)", "SS M64 N64 MN/MN (8 k-steps)", "SS M128 N64 B MN-major",
	                        "TS M128 N16", "SS M128 N128"};
	for (int mode = 0; mode < 6; ++mode) {
		int iters = 2000;
		bench_mma<<<148, 128, 49152>>>(mode, iters, dc);
		CK(cudaDeviceSynchronize());
		long long cyc;
		CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
		printf("MMA  %-30s : %.1f cycles per group (148 CTAs)\n", mnames[mode], (double)cyc / iters);
	}
	const char *enames[] = {"ld x64", "ld 2*x32", "cvt.relu x32", "st x32", "ld+cvt+st (layer epilogue)", "ld+cvt+sts smem+fence"};
	for (int warps : {4, 8, 16}) {
		for (int mode = 0; mode < 6; ++mode) {
			if (mode == 5 && warps > 16)
				continue;
			int iters = 2000;
			bench_epi<<<148, warps * 32, 65536>>>(mode, iters, dc, sink);
			CK(cudaDeviceSynchronize());
			long long cyc;
			CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
			printf("EPI  warps=%2d %-28s : %.1f cycles/iter  (=> %.1f cycles per 128x64 tile-layer per SM)\n", warps, enames[mode],
			       (double)cyc / iters, (double)cyc / iters / (warps / 4));
		}
	}
	return 0;
}
