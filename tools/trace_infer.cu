// trace_infer.cu -- development tool: runs nrc_infer_kernel built with -DNRC_TRACE and prints CTA 0's event timeline
// (events of thread 0 of each slot: 1 = before issue, 2 = issued, 3 = accumulator ready, 4 = loaded + converted,
// 5 = operand stored) plus a per-phase average over the steady state.
#include "../vknrc_b200/csrc/nrc_infer.cu"
#include <cstdio>
#include <vector>
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
	const int mode = argc > 1 ? atoi(argv[1]) : 0;
	const uint64_t n = 1920 * 1080;
	__half *w, *x, *y; float *rec;
	cudaMalloc(&w, 6 * 8192); cudaMalloc(&x, n * 128); cudaMalloc(&y, n * 6); cudaMalloc(&rec, n * 56);
	cudaMemset(w, 0, 6 * 8192); cudaMemset(x, 0, n * 128); cudaMemset(rec, 0, n * 56);
	CUtensorMap tw, ti; mk(&tw, w, 323, 64); mk(&ti, x, n, 128);
	nrc::InferParams p{}; p.n = n; p.in_mode = mode ? nrc::NRC_IN_UNPACKED : nrc::NRC_IN_ENCODED; p.out_mode = nrc::NRC_OUT_F16VEC3; p.out = y;
	p.in = rec; p.in_stride_bytes = 56;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int it = 0; it < 3; ++it) {
		cudaEventRecord(e0);
		nrc::launch_infer(p, tw, ti, 148, 0);
		cudaEventRecord(e1);
		printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
		float ms; cudaEventElapsedTime(&ms, e0, e1); printf("kernel %.1f us (with tracing)\n", ms * 1e3);
	}
	static uint2 tr[4][NRC_TRACE_CAP]; unsigned int cnt[4];
	cudaMemcpyFromSymbol(tr, g_nrc_trace, sizeof(tr)); cudaMemcpyFromSymbol(cnt, g_nrc_trace_n, sizeof(cnt));
	for (unsigned i = 200; i < cnt[0] && i < 280; ++i) {
		uint32_t tag = tr[0][i].x;
		printf("slot0 ev=%u layer=%u tile=%u t=%u (+%u)\n", tag >> 24, (tag >> 16) & 0xff, tag & 0xffff, tr[0][i].y - tr[0][0].y, tr[0][i].y - tr[0][i - 1].y);
	}
	printf("events %u, span %u cycles\n", cnt[0], tr[0][cnt[0] - 1].y - tr[0][0].y);
	// average phase durations (event a -> next event b) over slot 0
	double sum[8][8] = {}; long num[8][8] = {};
	for (unsigned i = 50; i + 1 < cnt[0]; ++i) {
		int a = (int)(tr[0][i].x >> 24), b = (int)(tr[0][i + 1].x >> 24);
		sum[a][b] += (double)(tr[0][i + 1].y - tr[0][i].y); num[a][b]++;
	}
	for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) if (num[a][b]) printf("phase %d->%d : avg %.0f cycles (n=%ld)\n", a, b, sum[a][b] / num[a][b], num[a][b]);
	return 0;
}
