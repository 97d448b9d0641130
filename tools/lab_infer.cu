// lab_infer.cu -- development tool: times nrc_infer_kernel (built from the product source with whatever -D switches
// are being compared) on the 1080p workload with CUDA events; no correctness check (tests/ does that).
#include "../vknrc_b200/csrc/nrc_infer.cu"
#include <cstdio>
#include <cstdlib>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void mk(CUtensorMap *tm, void *base, uint64_t rows, uint32_t box) {
	void *p; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	cuuint64_t gd[2] = {64, rows}, gs[1] = {128}; cuuint32_t bx[2] = {64, box}, es[2] = {1, 1};
	((PFN)p)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
	         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
// realistic switching activity: random fp16 inputs in [0,1), weights ~ U(-0.3,0.3) (He-normal-like scale), records in [0,1)
__global__ void fill_half(__half *p, uint64_t n, float lo, float hi, uint32_t seed) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t x = (uint32_t)i * 747796405u + seed; x = ((x >> ((x >> 28) + 4)) ^ x) * 277803737u; x = (x >> 22) ^ x;
		p[i] = __float2half(lo + (hi - lo) * (float)(x >> 8) * (1.0f / 16777216.0f));
	}
}
__global__ void fill_float(float *p, uint64_t n, uint32_t seed) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t x = (uint32_t)i * 747796405u + seed; x = ((x >> ((x >> 28) + 4)) ^ x) * 277803737u; x = (x >> 22) ^ x;
		p[i] = (float)(x >> 8) * (1.0f / 16777216.0f);
	}
}
int main(int argc, char **argv) {
	const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1920 * 1080;
	const bool random_data = !(argc > 2 && atoi(argv[2]) == 0);
	__half *w, *x, *y; float *rec;
	cudaMalloc(&w, 6 * 8192); cudaMalloc(&x, n * 128); cudaMalloc(&y, n * 6); cudaMalloc(&rec, n * 56);
	cudaMemset(w, 0, 6 * 8192); cudaMemset(x, 0x11, n * 128); cudaMemset(rec, 0x11, n * 56);
	if (random_data) {
		fill_half<<<1184, 256>>>(w, 6 * 4096, -0.3f, 0.3f, 1u); fill_half<<<1184, 256>>>(x, n * 64, 0.0f, 1.0f, 2u); fill_float<<<1184, 256>>>(rec, n * 14, 3u);
	}
	CUtensorMap tw, ti; mk(&tw, w, 323, 64); mk(&ti, x, n, 128);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int mode = 0; mode < 2; ++mode) {
		nrc::InferParams p{}; p.n = n; p.in_mode = mode ? nrc::NRC_IN_UNPACKED : nrc::NRC_IN_ENCODED; p.out_mode = nrc::NRC_OUT_F16VEC3; p.out = y;
		p.in = rec; p.in_stride_bytes = 56; p.clamp_output = 1;
		for (int it = 0; it < 5; ++it) nrc::launch_infer(p, tw, ti, 148, 0);
		cudaEventRecord(e0);
		const int iters = argc > 3 ? atoi(argv[3]) : 50;
		for (int it = 0; it < iters; ++it) nrc::launch_infer(p, tw, ti, 148, 0);
		cudaEventRecord(e1);
		cudaError_t e = cudaDeviceSynchronize();
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("%s n=%llu: %.2f us/launch  %.1f TFLOP/s (%s)\n", mode ? "unpacked " : "preencoded", (unsigned long long)n, ms * 1e3 / iters,
		       n * 41344.0 / (ms * 1e-3 / iters) / 1e12, cudaGetErrorString(e));
	}
	return 0;
}
