// trace_grad.cu -- development tool: runs nrc_gradient_kernel built with -DNRC_TRACE on one 16384-record batch and
// prints CTA 0 / thread 0's event timeline (tags: 1 entry, 2 setup done, 3 inputs ready, 0x1l fwd accumulator l ready,
// 0x2l fwd epilogue l stored, 0x3l dA accumulator ready, 0x4l delta stored, 5 tile loop done, 6 dW complete, 7 dW
// staged, 8 partial written, 9 grid barrier passed, 0x50-0x54 reduction (0x51 partials loaded, 0x58/0x52 tree, 0x55 gradient stored,
// 0x56 Adam done, 0x57 CTA barrier), 0x59 second grid barrier passed, 0x5A weights re-staged, 0x6k forward copy of layer k issued,
// 0x7l delta_l-1 copy stored + arrived, 0x8l / 0x9l drain of a finished dW accumulator started / done, 10 exit).
#include "../vknrc_b200/csrc/nrc_train.cu"
#include <cstdio>
#include <climits>
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
	const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 16384;
	__half *w; float *rec, *tgt, *partials;
	cudaMalloc(&w, 6 * 8192); cudaMalloc(&rec, n * 56); cudaMalloc(&tgt, n * 12); cudaMalloc(&partials, 148 * NRC_GRAD_STRIDE * 4);
	cudaMemset(w, 0, 6 * 8192); cudaMemset(rec, 0x11, n * 56); cudaMemset(tgt, 0, n * 12);
	{ // realistic values (He-normal-like weights, records and targets in [0, 1), optimizer moments of a run in progress): all-zero
	  // buffers send every IEEE division of the Adam step through its slow path and distort the timeline
		std::vector<__half> hw(6 * 4096); std::vector<float> hr(n * 14), ht(n * 3);
		uint32_t x = 12345u; auto rnd = [&]() { x = x * 1664525u + 1013904223u; return (float)(x >> 8) * (1.0f / 16777216.0f); };
		for (auto &v : hw) v = __float2half(0.6f * rnd() - 0.3f);
		for (auto &v : hr) v = rnd();
		for (auto &v : ht) v = rnd();
		cudaMemcpy(w, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(rec, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
		cudaMemcpy(tgt, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
	}
	CUtensorMap tw; mk(&tw, w, 323, 64);
	const int nb = argc > 2 ? atoi(argv[2]) : 1; // batches per launch (frame = 4)
	const bool encoded = argc > 3 && atoi(argv[3]) == 0; // argv[3]: 1 = 14-float records (default), 0 = pre-encoded inputs
	__half *enc = nullptr, *tgt16 = nullptr; CUtensorMap tin = tw;
	if (encoded) {
		cudaMalloc(&enc, n * 128); cudaMalloc(&tgt16, n * 6);
		std::vector<__half> he(n * 64), ht(n * 3);
		uint32_t x = 777u; auto rnd = [&]() { x = x * 1664525u + 1013904223u; return (float)(x >> 8) * (1.0f / 16777216.0f); };
		for (auto &v : he) v = __float2half(rnd());
		for (auto &v : ht) v = __float2half(rnd());
		cudaMemcpy(enc, he.data(), he.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(tgt16, ht.data(), ht.size() * 2, cudaMemcpyHostToDevice);
		mk(&tin, enc, n, 128);
	}
	NrcOptimizerEntry *entries; NrcOptimizerState *ost; uint32_t *sync; float *grads; __half *uw;
	cudaMalloc(&entries, 20672 * 16); cudaMalloc(&ost, 20); cudaMalloc(&sync, 32); cudaMalloc(&grads, NRC_GRAD_STRIDE * 4); cudaMalloc(&uw, 6 * 8192);
	cudaMemset(sync, 0, 32);
	{
		std::vector<NrcOptimizerEntry> he(20672); std::vector<__half> hw(20672);
		cudaMemcpy(hw.data(), w, 20672 * 2, cudaMemcpyDeviceToHost);
		for (int i = 0; i < 20672; ++i) he[i] = NrcOptimizerEntry{1e-3f * (float)((i % 7) - 3), 1e-6f * (float)(1 + i % 5), __half2float(hw[i]), __half2float(hw[i])};
		cudaMemcpy(entries, he.data(), he.size() * 16, cudaMemcpyHostToDevice);
	}
	const NrcOptimizerState st0{0u, 1.0f, 1.0f, 1.0f, 0.0f}; cudaMemcpy(ost, &st0, 20, cudaMemcpyHostToDevice);
	nrc::TrainParams tp{};
	for (int b = 0; b < nb; ++b) {
		nrc::GradParams &p = tp.batch[b];
		p.n = n; p.in_mode = nrc::NRC_IN_UNPACKED; p.loss_kind = nrc::NRC_LOSS_RELATIVE_L2_LUMINANCE; p.loss_scale = 1.0f;
		p.in = rec; p.in_stride_bytes = 56; p.target = tgt; p.target_stride_bytes = 12; p.partials = partials;
		if (encoded) { p.in_mode = nrc::NRC_IN_ENCODED; p.target = tgt16; p.target_stride_bytes = 6; p.target_is_f16 = 1; }
		tp.adam_mode[b] = b == nb - 1 ? 2 : 1;
	}
	tp.num_batches = nb; tp.gradients = grads; tp.limit = NRC_GRAD_STRIDE; tp.batch_cap = (uint32_t)n; tp.grid_bar = sync + 2;
	tp.adam.gradients = grads; tp.adam.entries = entries; tp.adam.opt_state = ost; tp.adam.done_counter = sync; tp.adam.weights = w; tp.adam.use_weights = uw;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int it = 0; it < 4; ++it) {
		cudaEventRecord(e0);
		cudaError_t le = nrc::launch_train(tp, tw, tin, 148, 0);
		cudaEventRecord(e1);
		cudaError_t e = cudaDeviceSynchronize();
		float ms; cudaEventElapsedTime(&ms, e0, e1); printf("train kernel (%d batch%s of %llu) %.1f us (%s / %s)\n", nb, nb > 1 ? "es" : "", (unsigned long long)n, ms * 1e3, cudaGetErrorString(le), cudaGetErrorString(e));
	}
	static uint2 tr[NRC_GTRACE_CAP]; unsigned int cnt;
	cudaMemcpyFromSymbol(tr, g_nrc_gtrace, sizeof(tr)); cudaMemcpyFromSymbol(&cnt, g_nrc_gtrace_n, sizeof(cnt));
	static uint2 it[NRC_GTRACE_CAP]; unsigned int icnt;
	cudaMemcpyFromSymbol(it, g_nrc_itrace, sizeof(it)); cudaMemcpyFromSymbol(&icnt, g_nrc_itrace_n, sizeof(icnt));
	static uint2 pt[NRC_GTRACE_CAP]; unsigned int pcnt;
	cudaMemcpyFromSymbol(pt, g_nrc_ptrace, sizeof(pt)); cudaMemcpyFromSymbol(&pcnt, g_nrc_ptrace_n, sizeof(pcnt));
	// merged by time stamp: epilogue thread (tags < 0x100), issuing thread (0x16k F wait done, 0x17k F issued, 0x18l B wait done,
	// 0x19l dA issued, 0x1Al delta in smem, 0x1Bl dW issued) and the first producer thread (0x200+t unit of tile t started, 0x210+t
	// encoded, 0x220+t stored and arrived)
	unsigned a = 0, b = 0, c = 0; uint32_t prev = tr[0].y;
	const unsigned lim = argc > 4 ? atoi(argv[4]) : 400;
	while ((a < cnt || b < icnt || c < pcnt) && a + b + c < lim) {
		const int32_t ta = a < cnt ? (int32_t)(tr[a].y - tr[0].y) : INT32_MAX, tb = b < icnt ? (int32_t)(it[b].y - tr[0].y) : INT32_MAX,
		              tc = c < pcnt ? (int32_t)(pt[c].y - tr[0].y) : INT32_MAX;
		const int who = ta <= tb && ta <= tc ? 0 : (tb <= tc ? 1 : 2);
		const uint2 e = who == 0 ? tr[a++] : who == 1 ? it[b++] : pt[c++];
		printf("%s ev=0x%03x t=%u (+%u)\n", who == 0 ? "epi  " : who == 1 ? "issue" : "prod ", e.x, e.y - tr[0].y, e.y - prev);
		prev = e.y;
	}
	return 0;
}
