// nrc_state.cu -- NrcState (the VkNRCState-shaped host object) and the C ABI declared in include/nrc_b200.h.
#include "nrc_state.hpp"

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/nrc_b200.h"

namespace nrc {

// ---------------------------------------------------------------------------------------------------- error plumbing
static thread_local std::string g_last_error;
static int set_error(int code, const std::string &what) {
	g_last_error = what;
	return code;
}
#define NRC_CUDA_TRY(expr, errsink)                                                                                   \
	do {                                                                                                               \
		cudaError_t e_ = (expr);                                                                                       \
		if (e_ != cudaSuccess)                                                                                         \
			return errsink(e_ == cudaErrorMemoryAllocation ? NRC_ERR_OUT_OF_MEMORY : NRC_ERR_CUDA,                     \
			               std::string(#expr) + ": " + cudaGetErrorString(e_));                                         \
	} while (0)

// ---------------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
	static PFN_encodeTiled fn = [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
			p = nullptr;
		return (PFN_encodeTiled)p;
	}();
	return fn;
}
// [rows][64] fp16 row-major tensor, box = box_rows x 64, 128-byte swizzle, out-of-bounds rows read as zero.
static int make_map(CUtensorMap *tm, const void *base, uint64_t rows, uint32_t box_rows, std::string *err) {
	PFN_encodeTiled fn = encode_fn();
	if (!fn) {
		*err = "cuTensorMapEncodeTiled is not available from this driver";
		return NRC_ERR_CUDA;
	}
	if (((uintptr_t)base & 15u) != 0) {
		*err = "fp16 matrix base address must be 16-byte aligned";
		return NRC_ERR_INVALID_ARGUMENT;
	}
	const cuuint64_t gdim[2] = {64, rows ? rows : 1};
	const cuuint64_t gstride[1] = {128};
	const cuuint32_t box[2] = {64, box_rows};
	const cuuint32_t estride[2] = {1, 1};
	CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		*err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
		return NRC_ERR_CUDA;
	}
	return NRC_OK;
}
int make_weight_tensor_map(CUtensorMap *tm, const void *d_weights, std::string *err) {
	return make_map(tm, d_weights, NRC_WEIGHT_ROWS, 64, err);
}
int make_input_tensor_map(CUtensorMap *tm, const void *d_inputs, uint64_t rows, std::string *err) {
	return make_map(tm, d_inputs, rows, NRC_TILE, err);
}
// A tensor map depends only on (base, rows, box): a renderer passes the same few buffers every frame, so the encoded
// descriptors are kept (cuTensorMapEncodeTiled costs ~1 us of host time per call, up to 9 per host-buffer inference).
int NrcState::cached_map(CUtensorMap *tm, const void *base, uint64_t rows, uint32_t box_rows, std::string *err) {
	for (MapEntry &e : m_maps)
		if (e.base == base && e.rows == rows && e.box == box_rows) {
			e.stamp = ++m_map_clock;
			*tm = e.map;
			return NRC_OK;
		}
	int rc = make_map(tm, base, rows, box_rows, err);
	if (rc != NRC_OK)
		return rc;
	MapEntry *slot = nullptr;
	if (m_maps.size() < kMapCacheEntries) {
		m_maps.emplace_back();
		slot = &m_maps.back();
	} else {
		slot = &m_maps[0];
		for (MapEntry &e : m_maps)
			if (e.stamp < slot->stamp)
				slot = &e;
	}
	slot->base = base, slot->rows = rows, slot->box = box_rows, slot->stamp = ++m_map_clock, slot->map = *tm;
	return NRC_OK;
}

static int check_device(int device, int *sms, std::string *err) {
	cudaDeviceProp prop;
	cudaError_t e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess) {
		*err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
		return NRC_ERR_CUDA;
	}
	if (prop.major != 10 || prop.minor != 0) { // the library holds one sm_100a cubin (tcgen05 / TMEM): no other device can load it
		*err = std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
		       "; this library only runs on sm_100 (B200)";
		return NRC_ERR_UNSUPPORTED_DEVICE;
	}
	*sms = prop.multiProcessorCount < (int)kMaxTrainGrid ? prop.multiProcessorCount : (int)kMaxTrainGrid;
	return NRC_OK;
}

// ---------------------------------------------------------------------------------------------------- NrcState
int NrcState::fail(int code, const std::string &what) {
	m_error_code = code;
	m_error = what;
	return set_error(code, what);
}

NrcState::NrcState(int device, Extent2D extent, uint64_t seed) : m_device(device), m_extent(extent), m_rng((uint32_t)seed) {
	std::string err;
	int rc = check_device(device, &m_sms, &err);
	if (rc != NRC_OK) {
		fail(rc, err);
		return;
	}
	auto alloc = [&](void **p, size_t bytes) {
		cudaError_t e = cudaMalloc(p, bytes);
		if (e != cudaSuccess) {
			fail(e == cudaErrorMemoryAllocation ? NRC_ERR_OUT_OF_MEMORY : NRC_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
			return false;
		}
		return cudaMemset(*p, 0, bytes) == cudaSuccess;
	};
	if (cudaSetDevice(device) != cudaSuccess) {
		fail(NRC_ERR_CUDA, "cudaSetDevice failed");
		return;
	}
	// weights are padded to whole 64-row TMA boxes so that no box of the weight tensor map leaves the allocation
	const size_t wbytes = (size_t)NRC_LAYERS * 64 * 64 * sizeof(__half);
	if (!alloc((void **)&m_weights, wbytes) || !alloc((void **)&m_use_weights, wbytes) ||
	    !alloc((void **)&m_optimizer_state, sizeof(NrcOptimizerState)) ||
	    !alloc((void **)&m_optimizer_entries, sizeof(NrcOptimizerEntry) * NRC_WEIGHT_COUNT) ||
	    !alloc((void **)&m_gradients, sizeof(float) * NRC_GRAD_STRIDE) ||
	    !alloc((void **)&m_partials, sizeof(float) * NRC_GRAD_STRIDE * gradient_max_partials(m_sms)) ||
	    !alloc((void **)&m_sync_words, 8 * sizeof(uint32_t)))
		return;
	m_ok = true;
	if (ResetMLPBuffers(seed) != NRC_OK)
		m_ok = false;
}

NrcState::~NrcState() {
	cudaFree(m_weights), cudaFree(m_use_weights), cudaFree(m_optimizer_state), cudaFree(m_optimizer_entries);
	cudaFree(m_gradients), cudaFree(m_partials), cudaFree(m_sync_words);
	cudaFree(m_stage_in), cudaFree(m_stage_out);
	if (m_stream_in) {
		cudaStreamDestroy(m_stream_in), cudaStreamDestroy(m_stream_out), cudaEventDestroy(m_ev_start), cudaEventDestroy(m_ev_out);
		for (int c = 0; c < kHostChunks; ++c)
			cudaEventDestroy(m_ev_in[c]), cudaEventDestroy(m_ev_done[c]);
	}
	CommShutdown();
}

int NrcState::CommInit(uint32_t rank, uint32_t world, cudaIpcMemHandle_t *out_handle) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	if (world < 1 || world > NRC_MAX_RANKS || rank >= world || !out_handle)
		return fail(NRC_ERR_INVALID_ARGUMENT, "CommInit: need rank < world <= 8 and a handle to fill");
	CommShutdown();
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	NRC_CUDA_TRY(cudaMalloc((void **)&m_comm_local, kCommBytes), sink);
	NRC_CUDA_TRY(cudaMemset(m_comm_local, 0, kCommBytes), sink);
	NRC_CUDA_TRY(cudaDeviceSynchronize(), sink);
	NRC_CUDA_TRY(cudaIpcGetMemHandle(out_handle, m_comm_local), sink);
	m_comm_rank = rank, m_comm_world = world, m_comm_owned = true;
	NRC_CUDA_TRY(cudaMemset(m_sync_words + 4, 0, 2 * sizeof(uint32_t)), sink); // exchange epochs restart with the fresh (zeroed) inboxes; error word cleared
	return NRC_OK;
}
int NrcState::CommConnect(const cudaIpcMemHandle_t *all_handles) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	if (!m_comm_local || !m_comm_owned || !all_handles)
		return fail(NRC_ERR_INVALID_ARGUMENT, "CommConnect: call CommInit first");
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	for (uint32_t r = 0; r < m_comm_world; ++r) {
		if (r == m_comm_rank) {
			m_comm_inbox[r] = m_comm_local;
			continue;
		}
		void *p = nullptr;
		NRC_CUDA_TRY(cudaIpcOpenMemHandle(&p, all_handles[r], cudaIpcMemLazyEnablePeerAccess), sink);
		m_comm_inbox[r] = (uint64_t *)p;
	}
	m_comm_connected = true;
	return NRC_OK;
}
// Caller-owned exchange buffers: any allocation every rank can address works - cudaDeviceEnablePeerAccess mappings in a
// single process that drives several GPUs, cuMemCreate / cuMulticastCreate, an NCCL window, torch symmetric memory.
// `multicast` (optional) maps all the inboxes through one NVSwitch multicast object: the push becomes one multimem.st.
int NrcState::CommAttach(uint32_t rank, uint32_t world, void *const *inboxes, void *multicast) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	if (world < 1 || world > NRC_MAX_RANKS || rank >= world || !inboxes)
		return fail(NRC_ERR_INVALID_ARGUMENT, "CommAttach: need rank < world <= 8 and world inbox pointers");
	for (uint32_t r = 0; r < world; ++r)
		if (!inboxes[r] || ((uintptr_t)inboxes[r] & 7u))
			return fail(NRC_ERR_INVALID_ARGUMENT, "CommAttach: every inbox must be a mapped, 8-byte aligned device pointer");
	if ((uintptr_t)multicast & 7u)
		return fail(NRC_ERR_INVALID_ARGUMENT, "CommAttach: the multicast pointer must be 8-byte aligned");
	CommShutdown();
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	for (uint32_t r = 0; r < world; ++r)
		m_comm_inbox[r] = (uint64_t *)inboxes[r];
	m_comm_local = m_comm_inbox[rank], m_comm_multicast = (uint64_t *)multicast;
	m_comm_rank = rank, m_comm_world = world, m_comm_owned = false;
	NRC_CUDA_TRY(cudaMemset(m_comm_local, 0, kCommBytes), sink); // (the caller synchronises the ranks after attaching, before the first training call)
	NRC_CUDA_TRY(cudaMemset(m_sync_words + 4, 0, 2 * sizeof(uint32_t)), sink);
	NRC_CUDA_TRY(cudaDeviceSynchronize(), sink);
	m_comm_connected = true;
	return NRC_OK;
}
int NrcState::CommShutdown() {
	if (m_comm_owned) {
		for (uint32_t r = 0; r < NRC_MAX_RANKS; ++r)
			if (m_comm_inbox[r] && m_comm_inbox[r] != m_comm_local)
				cudaIpcCloseMemHandle(m_comm_inbox[r]);
		if (m_comm_local)
			cudaFree(m_comm_local);
	}
	for (uint32_t r = 0; r < NRC_MAX_RANKS; ++r)
		m_comm_inbox[r] = nullptr;
	m_comm_local = m_comm_multicast = nullptr, m_comm_connected = false, m_comm_owned = false, m_comm_world = 1, m_comm_rank = 0;
	return NRC_OK;
}
// Synchronises `stream` and reports whether any exchange since the last set-up gave up waiting for a peer.
int NrcState::CommStatus(cudaStream_t stream) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	NRC_CUDA_TRY(cudaStreamSynchronize(stream), sink);
	uint32_t flag = 0;
	NRC_CUDA_TRY(cudaMemcpy(&flag, m_sync_words + 5, sizeof(flag), cudaMemcpyDeviceToHost), sink);
	if (flag)
		return fail(NRC_ERR_PEER_TIMEOUT, "a peer's gradient words did not arrive within the exchange timeout; the optimizer step of that batch was skipped "
		                                  "and the replicas may have diverged (re-synchronise the weights, then nrc_comm_init / nrc_comm_attach again)");
	return NRC_OK;
}

// src/VkNRCState.cpp:39-44 (He-normal, sigma = sqrt(2 / 64)) and :46-58 (fp32 master = ema = init, fp16 copy RNE)
int NrcState::ResetMLPBuffers(uint64_t seed) {
	std::mt19937 rng((uint32_t)seed);
	std::normal_distribution<float> norm{0, std::sqrt(2.0f / float(kNNWidth))};
	std::vector<float> w(NRC_WEIGHT_COUNT);
	for (auto &x : w)
		x = norm(rng);
	return upload_initial(w.data());
}
int NrcState::SetWeights(const float *fp32_weights) {
	if (!fp32_weights)
		return fail(NRC_ERR_INVALID_ARGUMENT, "SetWeights: null weights");
	return upload_initial(fp32_weights);
}
int NrcState::upload_initial(const float *w) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	std::vector<__half> h(NRC_WEIGHT_COUNT);
	std::vector<NrcOptimizerEntry> e(NRC_WEIGHT_COUNT);
	for (uint32_t i = 0; i < NRC_WEIGHT_COUNT; ++i) {
		h[i] = __float2half_rn(w[i]);
		e[i] = NrcOptimizerEntry{0.0f, 0.0f, w[i], w[i]};
	}
	const NrcOptimizerState st{0u, 1.0f, 1.0f, 1.0f, 0.0f}; // src/VkNRCState.cpp:50
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	// (blocking copies on the legacy stream do not order against the library's non-blocking streams: drain the device first,
	// so that no training launch in flight sees half-written weights)
	NRC_CUDA_TRY(cudaDeviceSynchronize(), sink);
	NRC_CUDA_TRY(cudaMemcpy(m_weights, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice), sink);
	NRC_CUDA_TRY(cudaMemcpy(m_use_weights, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice), sink);
	NRC_CUDA_TRY(cudaMemcpy(m_optimizer_entries, e.data(), e.size() * sizeof(NrcOptimizerEntry), cudaMemcpyHostToDevice), sink);
	NRC_CUDA_TRY(cudaMemcpy(m_optimizer_state, &st, sizeof(st), cudaMemcpyHostToDevice), sink);
	NRC_CUDA_TRY(cudaMemset(m_sync_words, 0, sizeof(uint32_t)), sink); // only the optimizer's "last CTA" counter: the grid-barrier counter
	                                                                 // is monotonic and base-relative, the exchange epochs restart only with fresh inboxes
	return NRC_OK;
}

int NrcState::Download(uint16_t *weights, uint16_t *use_weights, void *entries, void *state, float *gradients, cudaStream_t stream) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	NRC_CUDA_TRY(cudaStreamSynchronize(stream), sink);
	if (weights)
		NRC_CUDA_TRY(cudaMemcpy(weights, m_weights, NRC_WEIGHT_COUNT * 2, cudaMemcpyDeviceToHost), sink);
	if (use_weights)
		NRC_CUDA_TRY(cudaMemcpy(use_weights, m_use_weights, NRC_WEIGHT_COUNT * 2, cudaMemcpyDeviceToHost), sink);
	if (entries)
		NRC_CUDA_TRY(cudaMemcpy(entries, m_optimizer_entries, NRC_WEIGHT_COUNT * sizeof(NrcOptimizerEntry), cudaMemcpyDeviceToHost), sink);
	if (state)
		NRC_CUDA_TRY(cudaMemcpy(state, m_optimizer_state, sizeof(NrcOptimizerState), cudaMemcpyDeviceToHost), sink);
	if (gradients)
		NRC_CUDA_TRY(cudaMemcpy(gradients, m_gradients, NRC_GRAD_STRIDE * sizeof(float), cudaMemcpyDeviceToHost), sink);
	return NRC_OK;
}

int NrcState::Infer(InferParams p, const void *encoded_inputs, const __half *weights, cudaStream_t stream) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	if (p.n == 0)
		return NRC_OK;
	CUtensorMap tm_w, tm_in;
	std::string err;
	int rc = cached_map(&tm_w, weights, NRC_WEIGHT_ROWS, 64, &err);
	if (rc == NRC_OK)
		rc = p.in_mode == NRC_IN_ENCODED ? cached_map(&tm_in, encoded_inputs, p.n, NRC_TILE, &err) : (tm_in = tm_w, NRC_OK);
	if (rc != NRC_OK)
		return fail(rc, err);
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	NRC_CUDA_TRY(launch_infer(p, tm_w, tm_in, m_sms, stream), sink);
	return NRC_OK;
}

// Host-buffer inference: the queries are cut in chunks of whole tiles; chunk c's host -> device copy, MLP launch and
// device -> host copy run on three streams (both copy engines + the SMs busy at once), ordered after what `stream` holds at
// the call; `stream` completes when the last outputs are on the host. `launch_chunk(first, count, d_in, d_out)` enqueues the
// kernel of one chunk on `stream`.
template <class LaunchChunk>
int NrcState::host_pipeline(const void *h_in, uint32_t in_bytes, void *h_out, uint32_t out_bytes, uint64_t n, cudaStream_t stream, LaunchChunk launch_chunk) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	if (!m_stream_in) {
		NRC_CUDA_TRY(cudaStreamCreateWithFlags(&m_stream_in, cudaStreamNonBlocking), sink);
		NRC_CUDA_TRY(cudaStreamCreateWithFlags(&m_stream_out, cudaStreamNonBlocking), sink);
		NRC_CUDA_TRY(cudaEventCreateWithFlags(&m_ev_start, cudaEventDisableTiming), sink);
		NRC_CUDA_TRY(cudaEventCreateWithFlags(&m_ev_out, cudaEventDisableTiming), sink);
		for (int c = 0; c < kHostChunks; ++c) {
			NRC_CUDA_TRY(cudaEventCreateWithFlags(&m_ev_in[c], cudaEventDisableTiming), sink);
			NRC_CUDA_TRY(cudaEventCreateWithFlags(&m_ev_done[c], cudaEventDisableTiming), sink);
		}
	}
	if (n * in_bytes > m_stage_in_bytes || n * out_bytes > m_stage_out_bytes) { // (re-allocation synchronises the device: a one-time cost per size)
		cudaFree(m_stage_in), cudaFree(m_stage_out);
		m_stage_in = m_stage_out = nullptr, m_stage_in_bytes = m_stage_out_bytes = 0;
		m_maps.clear(); // (descriptors of the freed staging buffer must not outlive it)
		NRC_CUDA_TRY(cudaMalloc(&m_stage_in, n * in_bytes), sink);
		NRC_CUDA_TRY(cudaMalloc(&m_stage_out, n * out_bytes), sink);
		m_stage_in_bytes = n * in_bytes, m_stage_out_bytes = n * out_bytes;
	}
	// chunks of whole 128-query tiles. The host -> device copies are the long pole and everything else hides under them - except
	// what follows the LAST copy (that chunk's kernel and device -> host copy): shrinking chunks (16 16 12 8 6 3 2 1
	// sixty-fourths). Measured on a 1080p frame of 20-byte records (bare 41.5 MB copy: 0.75 ms): eight equal chunks 0.861 ms,
	// this table 0.852, five / four / three / two tapered chunks 0.854 / 0.866 / 0.894 / 1.08 ms (a chunk's device -> host
	// copy only overlaps the host -> device copies of LATER chunks)
	static constexpr uint32_t kCum64[] = {0, 16, 32, 44, 52, 58, 61, 63, 64};
	constexpr int kTaperChunks = (int)(sizeof(kCum64) / sizeof(kCum64[0])) - 1;
	static_assert(kTaperChunks >= 1 && kTaperChunks <= kHostChunks, "one event pair per chunk");
	const uint64_t tiles = (n + NRC_TILE - 1) / NRC_TILE;
	const bool tapered = tiles >= 64 * 8; // (small inputs: equal chunks, at most one per tile)
	const int chunks = tapered ? kTaperChunks : (int)(tiles < (uint64_t)kHostChunks ? tiles : (uint64_t)kHostChunks);
	NRC_CUDA_TRY(cudaEventRecord(m_ev_start, stream), sink);
	NRC_CUDA_TRY(cudaStreamWaitEvent(m_stream_in, m_ev_start, 0), sink);
	NRC_CUDA_TRY(cudaStreamWaitEvent(m_stream_out, m_ev_start, 0), sink);
	uint64_t first = 0;
	for (int c = 0; c < chunks; ++c) {
		const uint64_t last_tile = tapered ? tiles * kCum64[c + 1] / 64 : tiles * (uint64_t)(c + 1) / (uint64_t)chunks;
		const uint64_t end = last_tile * NRC_TILE < n ? last_tile * NRC_TILE : n, cnt = end - first;
		uint8_t *d_in = (uint8_t *)m_stage_in + first * in_bytes, *d_out = (uint8_t *)m_stage_out + first * out_bytes;
		NRC_CUDA_TRY(cudaMemcpyAsync(d_in, (const uint8_t *)h_in + first * in_bytes, cnt * in_bytes, cudaMemcpyHostToDevice, m_stream_in), sink);
		NRC_CUDA_TRY(cudaEventRecord(m_ev_in[c], m_stream_in), sink);
		NRC_CUDA_TRY(cudaStreamWaitEvent(stream, m_ev_in[c], 0), sink);
		if (int rc = launch_chunk(first, cnt, d_in, d_out); rc != NRC_OK)
			return rc;
		NRC_CUDA_TRY(cudaEventRecord(m_ev_done[c], stream), sink);
		NRC_CUDA_TRY(cudaStreamWaitEvent(m_stream_out, m_ev_done[c], 0), sink);
		NRC_CUDA_TRY(cudaMemcpyAsync((uint8_t *)h_out + first * out_bytes, d_out, cnt * out_bytes, cudaMemcpyDeviceToHost, m_stream_out), sink);
		first = end;
	}
	NRC_CUDA_TRY(cudaEventRecord(m_ev_out, m_stream_out), sink);
	NRC_CUDA_TRY(cudaStreamWaitEvent(stream, m_ev_out, 0), sink); // `stream` completes when the last outputs are on the host
	return NRC_OK;
}

int NrcState::InferEncodedHost(const void *h_inputs, void *h_outputs, uint64_t n, int clamp_output, const __half *weights, cudaStream_t stream) {
	if (n == 0)
		return NRC_OK;
	return host_pipeline(h_inputs, 128, h_outputs, 6, n, stream, [&](uint64_t, uint64_t cnt, void *d_in, void *d_out) {
		InferParams p{};
		p.n = cnt, p.in_mode = NRC_IN_ENCODED, p.out_mode = NRC_OUT_F16VEC3, p.clamp_output = clamp_output, p.out = d_out;
		return Infer(p, d_in, weights, stream);
	});
}

// The same pipeline for the reference's own query format: 20-byte NRCEvalRecords (or bare 16-byte PackedNRCInputs) in host
// memory -> UnpackNRCInput + encode + MLP on the device -> max(y, 0) as fp16 x 3 per query back in host memory.
int NrcState::InferPackedHost(const void *h_records, uint32_t stride_bytes, uint32_t input_offset, void *h_outputs, uint64_t n, const NrcScene &scene,
                              const __half *weights, cudaStream_t stream) {
	if (n == 0)
		return NRC_OK;
	return host_pipeline(h_records, stride_bytes, h_outputs, 6, n, stream, [&](uint64_t, uint64_t cnt, void *d_in, void *d_out) {
		InferParams p{};
		p.n = cnt, p.in_mode = NRC_IN_PACKED, p.out_mode = NRC_OUT_F16VEC3, p.clamp_output = 1;
		p.in = (const uint8_t *)d_in + input_offset, p.in_stride_bytes = stride_bytes, p.scene = scene, p.out = d_out;
		return Infer(p, nullptr, weights, stream);
	});
}

int NrcState::Train(TrainParams tp, const void *encoded_inputs, const __half *weights, cudaStream_t stream) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	if (tp.num_batches == 0 || tp.num_batches > NRC_TRAIN_BATCH_COUNT)
		return fail(NRC_ERR_INVALID_ARGUMENT, "Train: num_batches must be 1..4");
	const bool encoded = tp.batch[0].in_mode == NRC_IN_ENCODED;
	if (encoded && tp.num_batches != 1)
		return fail(NRC_ERR_INVALID_ARGUMENT, "Train: pre-encoded inputs are single-batch");
	CUtensorMap tm_w, tm_in;
	std::string err;
	int rc = cached_map(&tm_w, weights, NRC_WEIGHT_ROWS, 64, &err);
	if (rc == NRC_OK)
		rc = encoded ? cached_map(&tm_in, encoded_inputs, tp.batch[0].n, NRC_TILE, &err) : (tm_in = tm_w, NRC_OK);
	if (rc != NRC_OK)
		return fail(rc, err);
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	for (uint32_t b = 0; b < tp.num_batches; ++b)
		tp.batch[b].partials = m_partials;
	tp.grid_bar = m_sync_words + 2;
	tp.adam.gradients = tp.gradients, tp.adam.entries = m_optimizer_entries, tp.adam.opt_state = m_optimizer_state;
	tp.adam.done_counter = m_sync_words, tp.adam.weights = m_weights, tp.adam.use_weights = m_use_weights;
	tp.adam.use_ema = m_use_ema_weights ? 1 : 0;
	tp.comm = CommParams{};
	if (m_comm_connected && m_comm_world > 1 && !tp.accumulate) { // (the handle-less test-harness calls never exchange)
		tp.comm.rank = m_comm_rank, tp.comm.world = m_comm_world, tp.comm.epoch_word = m_sync_words + 4;
		tp.comm.error_word = m_sync_words + 5, tp.comm.spin_limit = m_comm_spin_limit, tp.comm.multicast = m_comm_multicast;
		for (uint32_t r = 0; r < m_comm_world; ++r)
			tp.comm.inbox[r] = m_comm_inbox[r];
	}
	NRC_CUDA_TRY(launch_train(tp, tm_w, tm_in, m_sms, stream), sink);
	return NRC_OK;
}

int NrcState::AdamStep(bool write_use_weights, cudaStream_t stream) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	AdamParams a{};
	a.gradients = m_gradients, a.entries = m_optimizer_entries, a.opt_state = m_optimizer_state, a.done_counter = m_sync_words;
	a.weights = m_weights, a.use_weights = write_use_weights ? m_use_weights : nullptr, a.use_ema = m_use_ema_weights ? 1 : 0;
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	NRC_CUDA_TRY(launch_adam(a, stream), sink);
	return NRC_OK;
}

int NrcState::SgdStep(float lr, float batch, cudaStream_t stream) {
	auto sink = [&](int c, const std::string &s) { return fail(c, s); };
	SgdParams s{};
	s.gradients = m_gradients, s.entries = m_optimizer_entries, s.weights = m_weights, s.lr = lr, s.batch = batch;
	NRC_CUDA_TRY(cudaSetDevice(m_device), sink);
	NRC_CUDA_TRY(launch_sgd(s, stream), sink);
	return NRC_OK;
}

} // namespace nrc

// ====================================================================================================== C ABI
using namespace nrc;
struct nrc_state_t {
	NrcState state;
	nrc_state_t(int device, Extent2D e, uint64_t seed) : state(device, e, seed) {}
};

static int bad_arg(const char *what) { return set_error(NRC_ERR_INVALID_ARGUMENT, what); }
#define NRC_REQUIRE(cond, msg)                                                                                         \
	do {                                                                                                               \
		if (!(cond))                                                                                                   \
			return bad_arg(msg);                                                                                       \
	} while (0)

// ---- handle-less test-harness kernels: a lazily created per-device scratch state supplies the partial buffers
static NrcState *scratch_state(int *rc) {
	static std::mutex mu;
	static std::vector<NrcState *> per_device;
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) {
		*rc = set_error(NRC_ERR_CUDA, "cudaGetDevice failed");
		return nullptr;
	}
	std::lock_guard<std::mutex> lock(mu);
	if ((int)per_device.size() <= dev)
		per_device.resize(dev + 1, nullptr);
	if (!per_device[dev]) {
		NrcState *s = new (std::nothrow) NrcState(dev, Extent2D{0, 0}, 0);
		if (!s || !s->ok()) {
			*rc = set_error(s ? s->error_code() : NRC_ERR_OUT_OF_MEMORY, s ? s->error() : "host allocation failed");
			delete s;
			return nullptr;
		}
		per_device[dev] = s;
	}
	return per_device[dev];
}

extern "C" {

const char *nrc_last_error(void) { return g_last_error.c_str(); }

uint64_t nrc_get_eval_record_buffer_size(uint32_t w, uint32_t h) { return NrcState::GetEvalRecordBufferSize({w, h}); }
uint64_t nrc_get_batch_train_record_buffer_size(void) { return NrcState::GetBatchTrainRecordBufferSize(); }
uint32_t nrc_get_train_batch_count(void) { return NrcState::GetTrainBatchCount(); }
uint32_t nrc_get_train_batch_size(void) { return NrcState::GetTrainBatchSize(); }
uint32_t nrc_get_weight_count(void) { return NrcState::GetWeightCount(); }
float nrc_get_default_train_probability(void) { return NrcState::GetDefaultTrainProbability(); }

int nrc_create(const nrc_config_t *config, int device, nrc_handle_t *out) {
	NRC_REQUIRE(config && out, "nrc_create: null argument");
	*out = nullptr;
	nrc_state_t *h = new (std::nothrow) nrc_state_t(device, Extent2D{config->extent_width, config->extent_height}, config->seed);
	if (!h)
		return set_error(NRC_ERR_OUT_OF_MEMORY, "nrc_create: host allocation failed");
	if (!h->state.ok()) {
		const int code = h->state.error_code();
		const std::string msg = h->state.error();
		delete h;
		return set_error(code ? code : NRC_ERR_CUDA, msg);
	}
	*out = h;
	return NRC_OK;
}
void nrc_destroy(nrc_handle_t h) { delete h; }

int nrc_reset_mlp_buffers(nrc_handle_t h, uint64_t seed) {
	NRC_REQUIRE(h, "null handle");
	return h->state.ResetMLPBuffers(seed);
}
int nrc_set_weights(nrc_handle_t h, const float *w) {
	NRC_REQUIRE(h, "null handle");
	return h->state.SetWeights(w);
}
void *nrc_get_weight_buffer(nrc_handle_t h) { return h ? h->state.GetWeightBuffer() : nullptr; }
void *nrc_get_use_weight_buffer(nrc_handle_t h) { return h ? h->state.GetUseWeightBuffer() : nullptr; }
void *nrc_get_optimizer_entry_buffer(nrc_handle_t h) { return h ? h->state.GetOptimizerEntryBuffer() : nullptr; }
void *nrc_get_optimizer_state_buffer(nrc_handle_t h) { return h ? h->state.GetOptimizerStateBuffer() : nullptr; }
void *nrc_get_gradient_buffer(nrc_handle_t h) { return h ? h->state.GetGradientBuffer() : nullptr; }
int nrc_download(nrc_handle_t h, uint16_t *weights, uint16_t *use_weights, void *entries, void *state, float *gradients, void *stream) {
	NRC_REQUIRE(h, "null handle");
	return h->state.Download(weights, use_weights, entries, state, gradients, (cudaStream_t)stream);
}
void nrc_set_use_ema_weights(nrc_handle_t h, int v) {
	if (h)
		h->state.SetUseEMAWeights(v != 0);
}
int nrc_is_use_ema_weights(nrc_handle_t h) { return h && h->state.IsUseEMAWeights(); }
void nrc_set_train_probability(nrc_handle_t h, float p) {
	if (h)
		h->state.SetTrainProbability(p);
}
float nrc_get_train_probability(nrc_handle_t h) { return h ? h->state.GetTrainProbability() : 0.0f; }
uint32_t nrc_next_frame(nrc_handle_t h) { return h ? h->state.NextFrame() : 0u; }
uint32_t nrc_get_seed(nrc_handle_t h) { return h ? h->state.GetSeed() : 0u; }
void nrc_set_prediction_capture(nrc_handle_t h, float *d) {
	if (h)
		h->state.SetPredictionCapture(d);
}

// ---- multi-GPU exchange set-up (one process per GPU; the 64-byte handles travel through the caller's own channel,
// e.g. torch.distributed.all_gather, MPI or a socket)
uint32_t nrc_comm_handle_bytes(void) { return (uint32_t)sizeof(cudaIpcMemHandle_t); }
int nrc_comm_init(nrc_handle_t h, uint32_t rank, uint32_t world, void *out_handle) {
	NRC_REQUIRE(h && out_handle, "nrc_comm_init: null argument");
	return h->state.CommInit(rank, world, (cudaIpcMemHandle_t *)out_handle);
}
int nrc_comm_connect(nrc_handle_t h, const void *all_handles) {
	NRC_REQUIRE(h && all_handles, "nrc_comm_connect: null argument");
	return h->state.CommConnect((const cudaIpcMemHandle_t *)all_handles);
}
int nrc_comm_shutdown(nrc_handle_t h) {
	NRC_REQUIRE(h, "null handle");
	return h->state.CommShutdown();
}
uint32_t nrc_comm_world(nrc_handle_t h) { return h ? h->state.comm_world() : 0u; }
uint64_t nrc_comm_buffer_bytes(void) { return (uint64_t)kCommBytes; }
int nrc_comm_attach(nrc_handle_t h, uint32_t rank, uint32_t world, void *const *d_inboxes, void *d_multicast) {
	NRC_REQUIRE(h && d_inboxes, "nrc_comm_attach: null argument");
	return h->state.CommAttach(rank, world, d_inboxes, d_multicast);
}
int nrc_comm_status(nrc_handle_t h, void *stream) {
	NRC_REQUIRE(h, "null handle");
	return h->state.CommStatus((cudaStream_t)stream);
}
void nrc_comm_set_timeout(nrc_handle_t h, uint32_t polls) {
	if (h)
		h->state.CommSetTimeout(polls);
}

int nrc_mlp_evaluate_encoded(const void *d_weights, const void *d_inputs, void *d_outputs, uint64_t n, void *stream) {
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(d_weights && d_inputs && d_outputs, "nrc_mlp_evaluate_encoded: null buffer");
	int rc = NRC_OK;
	NrcState *s = scratch_state(&rc);
	if (!s)
		return rc;
	InferParams p{};
	p.n = n, p.in_mode = NRC_IN_ENCODED, p.out_mode = NRC_OUT_F16VEC3, p.clamp_output = 0, p.out = d_outputs;
	return s->Infer(p, d_inputs, (const __half *)d_weights, (cudaStream_t)stream);
}

int nrc_mlp_gradient_encoded(const void *d_weights, float *d_dw, const void *d_inputs, const void *d_targets, uint64_t n, void *stream) {
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(d_weights && d_dw && d_inputs && d_targets, "nrc_mlp_gradient_encoded: null buffer");
	int rc = NRC_OK;
	NrcState *s = scratch_state(&rc);
	if (!s)
		return rc;
	TrainParams tp{};
	GradParams &p = tp.batch[0];
	p.n = n, p.in_mode = NRC_IN_ENCODED, p.loss_kind = NRC_LOSS_L2, p.loss_scale = 1.0f;
	p.target = d_targets, p.target_stride_bytes = 6, p.target_is_f16 = 1;
	tp.num_batches = 1, tp.gradients = d_dw, tp.accumulate = 1, tp.limit = NRC_WEIGHT_COUNT; // dw += (test/train_NV.comp)
	return s->Train(tp, d_inputs, (const __half *)d_weights, (cudaStream_t)stream);
}

int nrc_infer_encoded(nrc_handle_t h, const void *d_inputs, void *d_out, uint64_t n, int clamp_output, void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(d_inputs && d_out, "nrc_infer_encoded: null buffer");
	InferParams p{};
	p.n = n, p.in_mode = NRC_IN_ENCODED, p.out_mode = NRC_OUT_F16VEC3, p.clamp_output = clamp_output, p.out = d_out;
	return h->state.Infer(p, d_inputs, h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

int nrc_infer_encoded_host(nrc_handle_t h, const void *h_inputs, void *h_outputs, uint64_t n, int clamp_output, void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(h_inputs && h_outputs, "nrc_infer_encoded_host: null buffer");
	return h->state.InferEncodedHost(h_inputs, h_outputs, n, clamp_output, h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

int nrc_infer_eval_records_host(nrc_handle_t h, const void *h_eval_records, uint64_t n, const NrcScene *scene, void *h_outputs, void *stream);

int nrc_infer_unpacked(nrc_handle_t h, const void *d_records, uint32_t stride_bytes, const uint32_t *d_count, uint64_t max_count,
                       void *d_out, void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (max_count == 0)
		return NRC_OK;
	NRC_REQUIRE(d_records && d_out, "nrc_infer_unpacked: null buffer");
	NRC_REQUIRE(stride_bytes >= 56 && stride_bytes % 8 == 0 && ((uintptr_t)d_records & 7u) == 0, "nrc_infer_unpacked: records must be 8-byte aligned, stride >= 56 and a multiple of 8");
	InferParams p{};
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_UNPACKED, p.out_mode = NRC_OUT_F16VEC3, p.clamp_output = 1;
	p.in = d_records, p.in_stride_bytes = stride_bytes, p.out = d_out;
	return h->state.Infer(p, nullptr, h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

int nrc_infer_scatter_unpacked(nrc_handle_t h, const uint32_t *d_dst, uint32_t dst_stride_bytes, const void *d_records, uint32_t stride_bytes,
                               const uint32_t *d_count, uint64_t max_count, void *d_bias_factor_r, const void *d_factor_gb,
                               uint32_t image_pitch, void *const d_train_records[4], void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (max_count == 0)
		return NRC_OK;
	NRC_REQUIRE(d_dst && d_records, "nrc_infer_scatter_unpacked: null buffer");
	NRC_REQUIRE(dst_stride_bytes >= 4 && dst_stride_bytes % 4 == 0, "nrc_infer_scatter_unpacked: dst stride must be a multiple of 4");
	NRC_REQUIRE(stride_bytes >= 56 && stride_bytes % 8 == 0 && ((uintptr_t)d_records & 7u) == 0, "nrc_infer_scatter_unpacked: records must be 8-byte aligned, stride >= 56 and a multiple of 8");
	InferParams p{};
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_UNPACKED, p.out_mode = NRC_OUT_SCATTER, p.clamp_output = 1;
	p.in = d_records, p.in_stride_bytes = stride_bytes;
	p.dst = d_dst, p.dst_stride_u32 = dst_stride_bytes / 4;
	p.bias_factor_r = d_bias_factor_r, p.factor_gb = d_factor_gb, p.image_pitch = image_pitch;
	for (int b = 0; b < NRC_TRAIN_BATCH_COUNT; ++b)
		p.train_records[b] = d_train_records ? d_train_records[b] : nullptr;
	return h->state.Infer(p, nullptr, h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

// ---- the reference's record formats (NRCEvalRecord / NRCTrainRecord) + scene gather
static int check_scene(const NrcScene *sc) {
	NRC_REQUIRE(sc, "null scene");
	NRC_REQUIRE(sc->vertices && sc->vertex_indices && sc->texcoords && sc->texcoord_indices && sc->materials && sc->material_ids && sc->transforms,
	            "scene: a buffer pointer is null");
	NRC_REQUIRE(sc->texture_count == 0 || sc->textures, "scene: texture table is null");
	NRC_REQUIRE(((uintptr_t)sc->transforms & 15u) == 0 && ((uintptr_t)sc->materials & 15u) == 0 && ((uintptr_t)sc->texcoords & 7u) == 0,
	            "scene: transforms / materials must be 16-byte aligned, texcoords 8-byte aligned");
	NRC_REQUIRE(((uintptr_t)sc->prim_table & 63u) == 0, "scene: prim_table must be 64-byte aligned");
	return NRC_OK;
}

int nrc_infer(nrc_handle_t h, const void *d_eval_records, const uint32_t *d_count, uint64_t max_count, const NrcScene *scene,
              void *d_bias_factor_r, const void *d_factor_gb, uint32_t image_pitch, void *const d_train_records[4], void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (max_count == 0)
		return NRC_OK;
	NRC_REQUIRE(d_eval_records && ((uintptr_t)d_eval_records & 3u) == 0, "nrc_infer: eval records must be 4-byte aligned");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	InferParams p{};
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_PACKED, p.out_mode = NRC_OUT_SCATTER, p.clamp_output = 1;
	p.in = (const uint8_t *)d_eval_records + offsetof(NrcEvalRecord, packed_input), p.in_stride_bytes = sizeof(NrcEvalRecord), p.scene = *scene;
	p.dst = (const uint32_t *)d_eval_records, p.dst_stride_u32 = sizeof(NrcEvalRecord) / 4;
	p.bias_factor_r = d_bias_factor_r, p.factor_gb = d_factor_gb, p.image_pitch = image_pitch;
	for (int b = 0; b < NRC_TRAIN_BATCH_COUNT; ++b)
		p.train_records[b] = d_train_records ? d_train_records[b] : nullptr;
	return h->state.Infer(p, nullptr, h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

int nrc_infer_packed(nrc_handle_t h, const void *d_packed_inputs, uint32_t stride_bytes, const uint32_t *d_count, uint64_t max_count,
                     const NrcScene *scene, void *d_out, void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (max_count == 0)
		return NRC_OK;
	NRC_REQUIRE(d_packed_inputs && d_out, "nrc_infer_packed: null buffer");
	NRC_REQUIRE(stride_bytes >= 16 && stride_bytes % 4 == 0 && ((uintptr_t)d_packed_inputs & 3u) == 0,
	            "nrc_infer_packed: inputs must be 4-byte aligned, stride >= 16 and a multiple of 4");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	InferParams p{};
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_PACKED, p.out_mode = NRC_OUT_F16VEC3, p.clamp_output = 1;
	p.in = d_packed_inputs, p.in_stride_bytes = stride_bytes, p.scene = *scene, p.out = d_out;
	return h->state.Infer(p, nullptr, h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

int nrc_infer_eval_records_host(nrc_handle_t h, const void *h_eval_records, uint64_t n, const NrcScene *scene, void *h_outputs, void *stream) {
	NRC_REQUIRE(h, "null handle");
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(h_eval_records && h_outputs, "nrc_infer_eval_records_host: null buffer");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	return h->state.InferPackedHost(h_eval_records, sizeof(NrcEvalRecord), offsetof(NrcEvalRecord, packed_input), h_outputs, n, *scene,
	                                h->state.GetUseWeightBuffer(), (cudaStream_t)stream);
}

uint64_t nrc_scene_prim_table_bytes(uint32_t prim_count) { return (uint64_t)prim_count * sizeof(NrcPrimRow); }
int nrc_scene_build_prim_table(const NrcScene *scene, uint32_t prim_count, void *d_prim_table, void *stream) {
	if (prim_count == 0)
		return NRC_OK;
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	NRC_REQUIRE(d_prim_table && ((uintptr_t)d_prim_table & 63u) == 0, "nrc_scene_build_prim_table: the table must be 64-byte aligned device memory");
	cudaError_t e = launch_prim_table(*scene, prim_count, d_prim_table, (cudaStream_t)stream);
	if (e != cudaSuccess)
		return set_error(NRC_ERR_CUDA, std::string("nrc_scene_build_prim_table: ") + cudaGetErrorString(e));
	return NRC_OK;
}
int nrc_unpack_inputs(const void *d_packed_inputs, uint32_t stride_bytes, uint64_t n, const NrcScene *scene, float *d_unpacked14, void *stream) {
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(d_packed_inputs && d_unpacked14, "nrc_unpack_inputs: null buffer");
	NRC_REQUIRE(stride_bytes >= 16 && stride_bytes % 4 == 0 && ((uintptr_t)d_packed_inputs & 3u) == 0 && ((uintptr_t)d_unpacked14 & 7u) == 0,
	            "nrc_unpack_inputs: inputs 4-byte aligned with stride >= 16 (multiple of 4), outputs 8-byte aligned");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	NrcState *s = scratch_state(&rc); // (also the "is this an sm_100 device" check, done once per device)
	if (!s)
		return rc;
	cudaError_t e = launch_unpack(d_packed_inputs, stride_bytes, n, *scene, d_unpacked14, (cudaStream_t)stream);
	if (e != cudaSuccess)
		return set_error(NRC_ERR_CUDA, std::string("nrc_unpack_kernel: ") + cudaGetErrorString(e));
	return NRC_OK;
}

int nrc_encode_inputs(const void *d_inputs14, uint32_t stride_bytes, uint64_t n, void *d_encoded_f16x64, void *stream) {
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(d_inputs14 && d_encoded_f16x64, "nrc_encode_inputs: null buffer");
	NRC_REQUIRE(stride_bytes >= 56 && stride_bytes % 8 == 0 && ((uintptr_t)d_inputs14 & 7u) == 0 && ((uintptr_t)d_encoded_f16x64 & 15u) == 0,
	            "nrc_encode_inputs: records 8-byte aligned with stride >= 56 (multiple of 8), encoded rows 16-byte aligned");
	int rc = NRC_OK;
	NrcState *s = scratch_state(&rc); // (the "is this an sm_100 device" check, done once per device)
	if (!s)
		return rc;
	cudaError_t e = launch_encode(d_inputs14, NRC_IN_UNPACKED, stride_bytes, n, NrcScene{}, d_encoded_f16x64, (cudaStream_t)stream);
	if (e != cudaSuccess)
		return set_error(NRC_ERR_CUDA, std::string("nrc_encode_kernel: ") + cudaGetErrorString(e));
	return NRC_OK;
}
int nrc_encode_packed_inputs(const void *d_packed_inputs, uint32_t stride_bytes, uint64_t n, const NrcScene *scene, void *d_encoded_f16x64, void *stream) {
	if (n == 0)
		return NRC_OK;
	NRC_REQUIRE(d_packed_inputs && d_encoded_f16x64, "nrc_encode_packed_inputs: null buffer");
	NRC_REQUIRE(stride_bytes >= 16 && stride_bytes % 4 == 0 && ((uintptr_t)d_packed_inputs & 3u) == 0 && ((uintptr_t)d_encoded_f16x64 & 15u) == 0,
	            "nrc_encode_packed_inputs: inputs 4-byte aligned with stride >= 16 (multiple of 4), encoded rows 16-byte aligned");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	NrcState *s = scratch_state(&rc);
	if (!s)
		return rc;
	cudaError_t e = launch_encode(d_packed_inputs, NRC_IN_PACKED, stride_bytes, n, *scene, d_encoded_f16x64, (cudaStream_t)stream);
	if (e != cudaSuccess)
		return set_error(NRC_ERR_CUDA, std::string("nrc_encode_kernel: ") + cudaGetErrorString(e));
	return NRC_OK;
}

static void fill_record_batch(nrc_handle_t h, GradParams &p, const void *d_records, uint32_t *d_count, uint32_t max_count, const NrcScene *scene) {
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_PACKED, p.loss_kind = NRC_LOSS_RELATIVE_L2_LUMINANCE, p.loss_scale = NRC_LOSS_SCALE;
	p.in = (const uint8_t *)d_records + offsetof(NrcTrainRecord, packed_input), p.in_stride_bytes = sizeof(NrcTrainRecord), p.scene = *scene;
	p.target = d_records, p.target_stride_bytes = sizeof(NrcTrainRecord), p.target_is_f16 = 0; // target = bias (nrc_gradient.comp:31-33)
	p.y_out = h->state.GetPredictionCapture();
}
static int train_records_impl(nrc_handle_t h, const void *d_records, uint32_t *d_count, uint32_t max_count, const NrcScene *scene, void *stream,
                              int adam_mode) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(max_count == 0 || (d_records && ((uintptr_t)d_records & 3u) == 0), "training: train records must be 4-byte aligned");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	TrainParams tp{};
	fill_record_batch(h, tp.batch[0], d_records, d_count, max_count, scene);
	tp.num_batches = 1, tp.adam_mode[0] = adam_mode, tp.gradients = h->state.GetGradientBuffer(), tp.limit = NRC_GRAD_STRIDE, tp.batch_cap = max_count;
	return h->state.Train(tp, nullptr, h->state.GetWeightBuffer(), (cudaStream_t)stream);
}
int nrc_gradient(nrc_handle_t h, const void *d_records, uint32_t *d_count, uint32_t max_count, const NrcScene *scene, void *stream) {
	return train_records_impl(h, d_records, d_count, max_count, scene, stream, 0);
}
int nrc_train_batch(nrc_handle_t h, const void *d_records, uint32_t *d_count, uint32_t max_count, const NrcScene *scene, int write_use_weights,
                    void *stream) {
	return train_records_impl(h, d_records, d_count, max_count, scene, stream, write_use_weights ? 2 : 1);
}
int nrc_train_frame(nrc_handle_t h, void *const d_records[4], uint32_t *const d_counts[4], uint32_t max_count, const NrcScene *scene, void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(d_records, "nrc_train_frame: null buffer table");
	int rc = check_scene(scene);
	if (rc != NRC_OK)
		return rc;
	TrainParams tp{};
	for (int b = 0; b < NRC_TRAIN_BATCH_COUNT; ++b) {
		NRC_REQUIRE(max_count == 0 || (d_records[b] && ((uintptr_t)d_records[b] & 3u) == 0), "nrc_train_frame: train records must be 4-byte aligned");
		fill_record_batch(h, tp.batch[b], d_records[b], d_counts ? d_counts[b] : nullptr, max_count, scene);
		if (b != NRC_TRAIN_BATCH_COUNT - 1)
			tp.batch[b].y_out = nullptr; // the capture buffer holds one batch: keep the last batch's predictions
		tp.adam_mode[b] = b == NRC_TRAIN_BATCH_COUNT - 1 ? 2 : 1; // use_weights by the last batch only (NRCRenderGraph.cpp:66-68)
	}
	tp.num_batches = NRC_TRAIN_BATCH_COUNT, tp.gradients = h->state.GetGradientBuffer(), tp.limit = NRC_GRAD_STRIDE, tp.batch_cap = max_count;
	return h->state.Train(tp, nullptr, h->state.GetWeightBuffer(), (cudaStream_t)stream);
}

// ---- one frame of the render graph, NN part (src/rg/NRCRenderGraph.cpp): PreExecute's counter reset (:108-112), then - after
// the caller's record producer - the nn_inference_pass (:46-55) and the four nn_train_pass groups (:57-80)
int nrc_frame_begin(nrc_handle_t h, uint32_t *d_eval_count, uint32_t *const d_train_counts[4], void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(d_eval_count && d_train_counts, "nrc_frame_begin: null count pointer");
	if (cudaError_t e = cudaMemsetAsync(d_eval_count, 0, sizeof(uint32_t), (cudaStream_t)stream); e != cudaSuccess)
		return set_error(NRC_ERR_CUDA, std::string("nrc_frame_begin: ") + cudaGetErrorString(e));
	for (int b = 0; b < NRC_TRAIN_BATCH_COUNT; ++b) {
		NRC_REQUIRE(d_train_counts[b], "nrc_frame_begin: null batch count pointer");
		if (cudaError_t e = cudaMemsetAsync(d_train_counts[b], 0, sizeof(uint32_t), (cudaStream_t)stream); e != cudaSuccess)
			return set_error(NRC_ERR_CUDA, std::string("nrc_frame_begin: ") + cudaGetErrorString(e));
	}
	return NRC_OK;
}
int nrc_frame(nrc_handle_t h, const void *d_eval_records, const uint32_t *d_eval_count, uint64_t max_eval_count, const NrcScene *scene,
              void *d_bias_factor_r, const void *d_factor_gb, uint32_t image_pitch, void *const d_train_records[4],
              uint32_t *const d_train_counts[4], void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(d_train_records && d_train_counts, "nrc_frame: null buffer table");
	int rc = nrc_infer(h, d_eval_records, d_eval_count, max_eval_count, scene, d_bias_factor_r, d_factor_gb, image_pitch, d_train_records, stream);
	if (rc != NRC_OK)
		return rc;
	return nrc_train_frame(h, d_train_records, d_train_counts, NRC_TRAIN_BATCH_SIZE, scene, stream);
}

static int check_unpacked_args(const char *who, const void *d_inputs, uint32_t input_stride, const void *d_targets, uint32_t target_stride,
                               uint32_t max_count) {
	(void)who;
	NRC_REQUIRE(max_count == 0 || (d_inputs && d_targets), "training: null record / target buffer");
	NRC_REQUIRE(input_stride >= 56 && input_stride % 8 == 0 && ((uintptr_t)d_inputs & 7u) == 0,
	            "training: inputs must be 8-byte aligned, stride >= 56 and a multiple of 8");
	NRC_REQUIRE(target_stride >= 12 && target_stride % 4 == 0 && ((uintptr_t)d_targets & 3u) == 0,
	            "training: targets must be 4-byte aligned, stride >= 12 and a multiple of 4");
	return NRC_OK;
}
static void fill_unpacked_batch(nrc_handle_t h, GradParams &p, const void *d_inputs, uint32_t input_stride, const void *d_targets,
                                uint32_t target_stride, uint32_t *d_count, uint32_t max_count) {
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_UNPACKED, p.loss_kind = NRC_LOSS_RELATIVE_L2_LUMINANCE, p.loss_scale = NRC_LOSS_SCALE;
	p.in = d_inputs, p.in_stride_bytes = input_stride, p.target = d_targets, p.target_stride_bytes = target_stride, p.target_is_f16 = 0;
	p.y_out = h->state.GetPredictionCapture();
}
// adam_mode: 0 = gradient + reduction only, 1 = + Adam/EMA step, 2 = + use_weights
static int train_unpacked_impl(nrc_handle_t h, const void *d_inputs, uint32_t input_stride, const void *d_targets, uint32_t target_stride,
                               uint32_t *d_count, uint32_t max_count, void *stream, int adam_mode) {
	NRC_REQUIRE(h, "null handle");
	int rc = check_unpacked_args("train", d_inputs, input_stride, d_targets, target_stride, max_count);
	if (rc != NRC_OK)
		return rc;
	TrainParams tp{};
	fill_unpacked_batch(h, tp.batch[0], d_inputs, input_stride, d_targets, target_stride, d_count, max_count);
	tp.num_batches = 1, tp.adam_mode[0] = adam_mode, tp.gradients = h->state.GetGradientBuffer(), tp.limit = NRC_GRAD_STRIDE, tp.batch_cap = max_count;
	return h->state.Train(tp, nullptr, h->state.GetWeightBuffer(), (cudaStream_t)stream);
}

int nrc_gradient_unpacked(nrc_handle_t h, const void *d_inputs, uint32_t input_stride, const void *d_targets, uint32_t target_stride,
                          uint32_t *d_count, uint32_t max_count, void *stream) {
	return train_unpacked_impl(h, d_inputs, input_stride, d_targets, target_stride, d_count, max_count, stream, 0);
}

int nrc_gradient_encoded(nrc_handle_t h, const void *d_inputs, const void *d_targets, uint32_t *d_count, uint32_t max_count, int relative_loss,
                         void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(max_count == 0 || (d_inputs && d_targets), "nrc_gradient_encoded: null buffer");
	TrainParams tp{};
	GradParams &p = tp.batch[0];
	p.n = max_count, p.d_count = d_count, p.in_mode = NRC_IN_ENCODED, p.loss_kind = relative_loss ? NRC_LOSS_RELATIVE_L2_LUMINANCE : NRC_LOSS_L2;
	p.loss_scale = NRC_LOSS_SCALE, p.target = d_targets, p.target_stride_bytes = 6, p.target_is_f16 = 1;
	p.y_out = h->state.GetPredictionCapture();
	tp.num_batches = 1, tp.gradients = h->state.GetGradientBuffer(), tp.limit = NRC_GRAD_STRIDE, tp.batch_cap = max_count;
	return h->state.Train(tp, d_inputs, h->state.GetWeightBuffer(), (cudaStream_t)stream);
}

int nrc_adam_step(nrc_handle_t h, int write_use_weights, void *stream) {
	NRC_REQUIRE(h, "null handle");
	return h->state.AdamStep(write_use_weights != 0, (cudaStream_t)stream);
}

// ONE launch: gradient pass, grid barrier, deterministic reduction of the partials and the Adam/EMA step
int nrc_train_batch_unpacked(nrc_handle_t h, const void *d_inputs, uint32_t input_stride, const void *d_targets, uint32_t target_stride,
                             uint32_t *d_count, uint32_t max_count, int write_use_weights, void *stream) {
	return train_unpacked_impl(h, d_inputs, input_stride, d_targets, target_stride, d_count, max_count, stream, write_use_weights ? 2 : 1);
}

// ONE launch for the whole frame: the four dependent batches of src/rg/NRCRenderGraph.cpp:57-70, use_weights published
// by the last one only (:66-68)
int nrc_train_frame_unpacked(nrc_handle_t h, const void *const d_inputs[4], uint32_t input_stride, const void *const d_targets[4],
                             uint32_t target_stride, uint32_t *const d_counts[4], uint32_t max_count, void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(d_inputs && d_targets, "nrc_train_frame_unpacked: null buffer table");
	TrainParams tp{};
	for (int b = 0; b < NRC_TRAIN_BATCH_COUNT; ++b) {
		int rc = check_unpacked_args("frame", d_inputs[b], input_stride, d_targets[b], target_stride, max_count);
		if (rc != NRC_OK)
			return rc;
		fill_unpacked_batch(h, tp.batch[b], d_inputs[b], input_stride, d_targets[b], target_stride, d_counts ? d_counts[b] : nullptr, max_count);
		if (tp.batch[b].y_out) // the capture buffer holds one batch: keep the last batch's predictions
			tp.batch[b].y_out = b == NRC_TRAIN_BATCH_COUNT - 1 ? tp.batch[b].y_out : nullptr;
		tp.adam_mode[b] = b == NRC_TRAIN_BATCH_COUNT - 1 ? 2 : 1;
	}
	tp.num_batches = NRC_TRAIN_BATCH_COUNT, tp.gradients = h->state.GetGradientBuffer(), tp.limit = NRC_GRAD_STRIDE, tp.batch_cap = max_count;
	return h->state.Train(tp, nullptr, h->state.GetWeightBuffer(), (cudaStream_t)stream);
}

int nrc_image_train_step(nrc_handle_t h, const void *d_image_rgba8, uint32_t image_w, uint32_t image_h, uint32_t seed_x, uint32_t seed_y,
                         uint32_t batch, float lr, void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(d_image_rgba8 && image_w && image_h && batch, "nrc_image_train_step: bad image or batch");
	TrainParams tp{};
	GradParams &p = tp.batch[0];
	p.n = batch, p.in_mode = NRC_IN_IMAGE_RANDOM, p.loss_kind = NRC_LOSS_L2, p.loss_scale = 1.0f;
	p.seed_x = seed_x, p.seed_y = seed_y, p.image_rgba8 = (const uint8_t *)d_image_rgba8, p.image_w = image_w, p.image_h = image_h;
	p.y_out = h->state.GetPredictionCapture();
	tp.num_batches = 1, tp.gradients = h->state.GetGradientBuffer(), tp.limit = NRC_GRAD_STRIDE;
	int rc = h->state.Train(tp, nullptr, h->state.GetWeightBuffer(), (cudaStream_t)stream);
	if (rc != NRC_OK)
		return rc;
	return h->state.SgdStep(lr, (float)batch, (cudaStream_t)stream);
}

int nrc_image_infer(nrc_handle_t h, void *d_out_rgba8, uint32_t width, void *stream) {
	NRC_REQUIRE(h, "null handle");
	NRC_REQUIRE(d_out_rgba8 && width, "nrc_image_infer: bad output");
	InferParams p{};
	p.n = (uint64_t)width * width, p.in_mode = NRC_IN_IMAGE_GRID, p.out_mode = NRC_OUT_RGBA8, p.image_width = width, p.out = d_out_rgba8;
	return h->state.Infer(p, nullptr, h->state.GetWeightBuffer(), (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Vulkan interop. The reference allocates the record / count / weight / image buffers as VkBuffers /
// VkImages (src/rg/NRCRenderGraph.cpp:139-175, src/VkNRCState.cpp:59-88); a renderer that keeps doing so exports each
// VkDeviceMemory as an opaque fd (VK_KHR_external_memory_fd) and its timeline semaphore (VK_KHR_external_semaphore_fd)
// and hands the fds over here. On success the fd is owned by the CUDA driver (do not close it), as CUDA specifies.
// ------------------------------------------------------------------------------------------------------------------
struct nrc_external_memory_t {
	int device;
	cudaExternalMemory_t mem;
	uint64_t size;
};
struct nrc_external_semaphore_t {
	int device;
	cudaExternalSemaphore_t sem;
};
static int cuda_fail(const char *where, cudaError_t e) {
	cudaGetLastError(); // (do not leave the error sticky for the caller's next runtime call)
	return set_error(e == cudaErrorMemoryAllocation ? NRC_ERR_OUT_OF_MEMORY : NRC_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

int nrc_import_vulkan_memory_fd(int device, int fd, uint64_t allocation_size, int dedicated, nrc_external_memory_handle_t *out) {
	NRC_REQUIRE(out, "nrc_import_vulkan_memory_fd: null output");
	*out = nullptr;
	NRC_REQUIRE(fd >= 0 && allocation_size > 0, "nrc_import_vulkan_memory_fd: bad fd or size");
	if (cudaError_t e = cudaSetDevice(device); e != cudaSuccess)
		return cuda_fail("nrc_import_vulkan_memory_fd: cudaSetDevice", e);
	cudaExternalMemoryHandleDesc d{};
	d.type = cudaExternalMemoryHandleTypeOpaqueFd, d.handle.fd = fd, d.size = allocation_size;
	d.flags = dedicated ? cudaExternalMemoryDedicated : 0; // VkMemoryDedicatedAllocateInfo allocations must say so
	nrc_external_memory_t *m = new (std::nothrow) nrc_external_memory_t{device, nullptr, allocation_size};
	if (!m)
		return set_error(NRC_ERR_OUT_OF_MEMORY, "nrc_import_vulkan_memory_fd: host allocation failed");
	if (cudaError_t e = cudaImportExternalMemory(&m->mem, &d); e != cudaSuccess) {
		delete m;
		return cuda_fail("cudaImportExternalMemory", e);
	}
	*out = m;
	return NRC_OK;
}

int nrc_external_memory_map_buffer(nrc_external_memory_handle_t m, uint64_t offset, uint64_t size, void **d_ptr) {
	NRC_REQUIRE(m && d_ptr, "nrc_external_memory_map_buffer: null argument");
	*d_ptr = nullptr;
	NRC_REQUIRE(size > 0 && offset <= m->size && size <= m->size - offset, "nrc_external_memory_map_buffer: range outside the allocation");
	if (cudaError_t e = cudaSetDevice(m->device); e != cudaSuccess)
		return cuda_fail("nrc_external_memory_map_buffer: cudaSetDevice", e);
	cudaExternalMemoryBufferDesc b{};
	b.offset = offset, b.size = size;
	if (cudaError_t e = cudaExternalMemoryGetMappedBuffer(d_ptr, m->mem, &b); e != cudaSuccess)
		return cuda_fail("cudaExternalMemoryGetMappedBuffer", e);
	return NRC_OK;
}

// every pointer mapped from `m` must have been cudaFree'd by the caller before (CUDA's rule for mapped buffers)
int nrc_external_memory_release(nrc_external_memory_handle_t m) {
	if (!m)
		return NRC_OK;
	cudaSetDevice(m->device);
	const cudaError_t e = cudaDestroyExternalMemory(m->mem);
	delete m;
	return e == cudaSuccess ? NRC_OK : cuda_fail("cudaDestroyExternalMemory", e);
}

int nrc_import_vulkan_timeline_semaphore_fd(int device, int fd, nrc_external_semaphore_handle_t *out) {
	NRC_REQUIRE(out, "nrc_import_vulkan_timeline_semaphore_fd: null output");
	*out = nullptr;
	NRC_REQUIRE(fd >= 0, "nrc_import_vulkan_timeline_semaphore_fd: bad fd");
	if (cudaError_t e = cudaSetDevice(device); e != cudaSuccess)
		return cuda_fail("nrc_import_vulkan_timeline_semaphore_fd: cudaSetDevice", e);
	cudaExternalSemaphoreHandleDesc d{};
	d.type = cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd, d.handle.fd = fd;
	nrc_external_semaphore_t *s = new (std::nothrow) nrc_external_semaphore_t{device, nullptr};
	if (!s)
		return set_error(NRC_ERR_OUT_OF_MEMORY, "nrc_import_vulkan_timeline_semaphore_fd: host allocation failed");
	if (cudaError_t e = cudaImportExternalSemaphore(&s->sem, &d); e != cudaSuccess) {
		delete s;
		return cuda_fail("cudaImportExternalSemaphore", e);
	}
	*out = s;
	return NRC_OK;
}

// the renderer signals `value` after the path tracer wrote the frame's records (src/rg/NRCRenderGraph.cpp:60-70 order)
int nrc_external_semaphore_wait(nrc_external_semaphore_handle_t s, uint64_t value, void *stream) {
	NRC_REQUIRE(s, "nrc_external_semaphore_wait: null semaphore");
	cudaExternalSemaphoreWaitParams p{};
	p.params.fence.value = value;
	if (cudaError_t e = cudaWaitExternalSemaphoresAsync(&s->sem, &p, 1, (cudaStream_t)stream); e != cudaSuccess)
		return cuda_fail("cudaWaitExternalSemaphoresAsync", e);
	return NRC_OK;
}

// ... and waits for `value` before the screen pass reads the composited image / the next frame reuses the records
int nrc_external_semaphore_signal(nrc_external_semaphore_handle_t s, uint64_t value, void *stream) {
	NRC_REQUIRE(s, "nrc_external_semaphore_signal: null semaphore");
	cudaExternalSemaphoreSignalParams p{};
	p.params.fence.value = value;
	if (cudaError_t e = cudaSignalExternalSemaphoresAsync(&s->sem, &p, 1, (cudaStream_t)stream); e != cudaSuccess)
		return cuda_fail("cudaSignalExternalSemaphoresAsync", e);
	return NRC_OK;
}

int nrc_external_semaphore_release(nrc_external_semaphore_handle_t s) {
	if (!s)
		return NRC_OK;
	cudaSetDevice(s->device);
	const cudaError_t e = cudaDestroyExternalSemaphore(s->sem);
	delete s;
	return e == cudaSuccess ? NRC_OK : cuda_fail("cudaDestroyExternalSemaphore", e);
}

} // extern "C"
