// trace_infer.cu -- development tool: runs nrc_infer_kernel built with -DNRC_TRACE and prints CTA 0's event timeline.
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
int main() {
	const uint64_t n = 1920 * 1080;
	__half *w, *x, *y;
	cudaMalloc(&w, 6 * 8192); cudaMalloc(&x, n * 128); cudaMalloc(&y, n * 6);
	cudaMemset(w, 0, 6 * 8192); cudaMemset(x, 0, n * 128);
	CUtensorMap tw, ti; mk(&tw, w, 323, 64); mk(&ti, x, n, 128);
	nrc::InferParams p{}; p.n = n; p.in_mode = nrc::NRC_IN_ENCODED; p.out_mode = nrc::NRC_OUT_F16VEC3; p.out = y;
	for (int it = 0; it < 2; ++it) {
		unsigned int z[4] = {0, 0, 0, 0};
		cudaMemcpyToSymbol(g_nrc_trace_n, z, sizeof(z));
		nrc::launch_infer(p, tw, ti, 148, 0);
		printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	}
	static unsigned long long tr[4][4096]; unsigned int cnt[4];
	cudaMemcpyFromSymbol(tr, g_nrc_trace, sizeof(tr)); cudaMemcpyFromSymbol(cnt, g_nrc_trace_n, sizeof(cnt));
	unsigned long long t0 = tr[0][1];
	for (int who = 0; who < 2; ++who)
		for (unsigned i = 0; i < cnt[who] && i < 400; ++i) {
			unsigned long long tag = tr[who][2 * i], t = tr[who][2 * i + 1];
			printf("who=%d ev=%llu layer=%llu tile=%llu t=%lld\n", who, tag >> 24, (tag >> 16) & 0xff, tag & 0xffff, (long long)(t - t0));
		}
	return 0;
}
