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
int main(int argc, char **argv) {
	const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 1920 * 1080;
	__half *w, *x, *y; float *rec;
	cudaMalloc(&w, 6 * 8192); cudaMalloc(&x, n * 128); cudaMalloc(&y, n * 6); cudaMalloc(&rec, n * 56);
	cudaMemset(w, 0, 6 * 8192); cudaMemset(x, 0x11, n * 128); cudaMemset(rec, 0x11, n * 56);
	CUtensorMap tw, ti; mk(&tw, w, 323, 64); mk(&ti, x, n, 128);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int mode = 0; mode < 2; ++mode) {
		nrc::InferParams p{}; p.n = n; p.in_mode = mode ? nrc::NRC_IN_UNPACKED : nrc::NRC_IN_ENCODED; p.out_mode = nrc::NRC_OUT_F16VEC3; p.out = y;
		p.in = rec; p.in_stride_bytes = 56; p.clamp_output = 1;
		for (int it = 0; it < 5; ++it) nrc::launch_infer(p, tw, ti, 148, 0);
		cudaEventRecord(e0);
		const int iters = 50;
		for (int it = 0; it < iters; ++it) nrc::launch_infer(p, tw, ti, 148, 0);
		cudaEventRecord(e1);
		cudaError_t e = cudaDeviceSynchronize();
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("%s n=%llu: %.2f us/launch  %.1f TFLOP/s (%s)\n", mode ? "unpacked " : "preencoded", (unsigned long long)n, ms * 1e3 / iters,
		       n * 41344.0 / (ms * 1e-3 / iters) / 1e12, cudaGetErrorString(e));
	}
	return 0;
}
