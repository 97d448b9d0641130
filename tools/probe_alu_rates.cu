// probe_alu_rates.cu -- issue rates of the instruction forms the record-mode producers are made of (development tool).
// One CTA of W warps per SM; every warp runs `iters` iterations of 16 independent chains of one instruction form; prints
// warp-instructions per cycle per SM sub-partition for W = 4 (one warp per scheduler) and W = 16 (four per scheduler).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

enum { FFMA_RRR, FFMA_IMM, FFMA2, FADD_RR, FMUL_RR, FMUL2, F2FP, FRND, MIX_FMA_ALU, MIX_FMA_F2FP, NFORMS };
static const char *kNames[NFORMS] = {"FFMA r,r,r", "FFMA r,imm,r", "FFMA2 (2 fma)", "FADD r,r", "FMUL r,r", "FMUL2", "F2FP.PACK_AB", "FRND.FLOOR + FFMA",
                                     "FFMA + LOP3 pairs", "FFMA + F2FP pairs"};

template <int FORM> __global__ void __launch_bounds__(1024, 1) k(int iters, float a, float b, long long *cycles, float *sink) {
	float x[16];
	float2 y[8];
	uint32_t z[16];
#pragma unroll
	for (int i = 0; i < 16; ++i)
		x[i] = a + (float)i + (float)threadIdx.x, z[i] = threadIdx.x * 17u + i;
#pragma unroll
	for (int i = 0; i < 8; ++i)
		y[i] = make_float2(x[2 * i], x[2 * i + 1]);
	__syncthreads();
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			if (FORM == FFMA_RRR) x[i] = fmaf(x[i], a, b);
			if (FORM == FFMA_IMM) x[i] = fmaf(x[i], 1.0009765625f, b);
			if (FORM == FADD_RR) x[i] = x[i] + b;
			if (FORM == FMUL_RR) x[i] = x[i] * a;
			if (FORM == FRND) x[i] = floorf(x[i]) + 0.0f * b;
			if (FORM == F2FP) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(z[i]) : "f"(x[i]), "f"(__uint_as_float(z[i])));
			if (FORM == MIX_FMA_ALU) { x[i] = fmaf(x[i], a, b); z[i] = (z[i] & 0x7fffffffu) ^ (uint32_t)it; }
			if (FORM == MIX_FMA_F2FP) { x[i] = fmaf(x[i], a, b); asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(z[i]) : "f"(a), "f"(__uint_as_float(z[i]))); }
		}
		if (FORM == FFMA2 || FORM == FMUL2) {
#pragma unroll
			for (int i = 0; i < 8; ++i)
				y[i] = FORM == FFMA2 ? __ffma2_rn(y[i], make_float2(a, a), make_float2(b, b)) : __fmul2_rn(y[i], make_float2(a, a));
#pragma unroll
			for (int i = 0; i < 8; ++i)
				y[i] = FORM == FFMA2 ? __ffma2_rn(y[i], make_float2(a, a), make_float2(b, b)) : __fmul2_rn(y[i], make_float2(a, a));
		}
	}
	const long long t1 = clock64();
	float s = 0.0f;
#pragma unroll
	for (int i = 0; i < 16; ++i)
		s += x[i] + __uint_as_float(z[i]);
#pragma unroll
	for (int i = 0; i < 8; ++i)
		s += y[i].x + y[i].y;
	if (s == 123.456f)
		*sink = s;
	if (threadIdx.x == 0 && blockIdx.x == 0)
		*cycles = t1 - t0;
}

template <int FORM> static void run(long long *d_cycles, float *d_sink) {
	const int iters = 4096;
	for (int warps : {4, 16}) {
		k<FORM><<<148, warps * 32>>>(iters, 1.0001f, 0.5f, d_cycles, d_sink);
		k<FORM><<<148, warps * 32>>>(iters, 1.0001f, 0.5f, d_cycles, d_sink);
		long long c = 0;
		cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
		const double per_iter = (FORM == MIX_FMA_ALU || FORM == MIX_FMA_F2FP || FORM == FRND) ? 32.0 : 16.0; // warp instructions per warp per iteration (FRND: + 1 FFMA each)
		const double ipc = per_iter * iters * (warps / 4.0) / (double)c;
		printf("%-22s %2d warps/SM: %6.3f warp-instr / cycle / SMSP  (%lld cycles)\n", kNames[FORM], warps, ipc, c);
	}
}

int main() {
	long long *d_cycles;
	float *d_sink;
	cudaMalloc(&d_cycles, 8), cudaMalloc(&d_sink, 4);
	run<FFMA_RRR>(d_cycles, d_sink), run<FFMA_IMM>(d_cycles, d_sink), run<FFMA2>(d_cycles, d_sink), run<FADD_RR>(d_cycles, d_sink);
	run<FMUL_RR>(d_cycles, d_sink), run<FMUL2>(d_cycles, d_sink), run<F2FP>(d_cycles, d_sink), run<FRND>(d_cycles, d_sink);
	run<MIX_FMA_ALU>(d_cycles, d_sink), run<MIX_FMA_F2FP>(d_cycles, d_sink);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
