// nrc_state.hpp -- host object of the NRC hot path, shaped like the reference's VkNRCState
// (src/VkNRCState.hpp:17-90): it owns the persistent MLP buffers (weights, use_weights, optimizer state/entries),
// the flags the UI toggles, and exposes Evaluate/Train entry points that enqueue the sm_100a kernels on the caller's
// stream. Per-frame record / count / image buffers stay caller-owned, as in the reference's render graph
// (src/rg/NRCRenderGraph.cpp:139-175).
#pragma once
#include <random>
#include <string>
#include <vector>

#include "nrc_kernels.h"

namespace nrc {

struct Extent2D {
	uint32_t width, height;
};

class NrcState final {
public:
	static constexpr float kDefaultTrainProbability = 0.03f; // src/VkNRCState.hpp:22
	static constexpr uint32_t kNNHiddenLayers = NRC_HIDDEN_LAYERS, kNNWidth = NRC_WIDTH, kNNOutWidth = NRC_OUT_WIDTH,
	                          kTrainBatchSize = NRC_TRAIN_BATCH_SIZE, kTrainBatchCount = NRC_TRAIN_BATCH_COUNT;
	static constexpr uint32_t kNNWeighCount = NRC_WEIGHT_COUNT;

	// throws nothing: on failure `ok()` is false and `error()` explains why
	NrcState(int device, Extent2D extent, uint64_t seed);
	~NrcState();
	NrcState(const NrcState &) = delete;
	NrcState &operator=(const NrcState &) = delete;

	bool ok() const { return m_ok; }
	int error_code() const { return m_error_code; }
	const std::string &error() const { return m_error; }

	// ---- getters, as VkNRCState::Get*Buffer (src/VkNRCState.hpp:43-46)
	__half *GetWeightBuffer() const { return m_weights; }
	__half *GetUseWeightBuffer() const { return m_use_weights; }
	NrcOptimizerEntry *GetOptimizerEntryBuffer() const { return m_optimizer_entries; }
	NrcOptimizerState *GetOptimizerStateBuffer() const { return m_optimizer_state; }
	float *GetGradientBuffer() const { return m_gradients; }

	bool IsUseEMAWeights() const { return m_use_ema_weights; }
	void SetUseEMAWeights(bool v) { m_use_ema_weights = v; }
	float GetTrainProbability() const { return m_train_probability; }
	void SetTrainProbability(float p) { m_train_probability = p; }
	uint32_t GetSeed() const { return m_seed; }
	uint32_t NextFrame() { // src/VkNRCState.hpp:74-78
		m_seed = std::uniform_int_distribution<uint32_t>{0}(m_rng);
		return m_seed;
	}

	// ---- statics (src/VkNRCState.hpp:83-89, src/VkNRCState.cpp:34-37)
	static uint64_t GetEvalRecordBufferSize(Extent2D extent) {
		return ((uint64_t)extent.width * extent.height + (uint64_t)kTrainBatchSize * kTrainBatchCount) * sizeof(NrcEvalRecord);
	}
	static uint64_t GetBatchTrainRecordBufferSize() { return (uint64_t)kTrainBatchSize * sizeof(NrcTrainRecord); }
	static constexpr uint32_t GetTrainBatchCount() { return kTrainBatchCount; }
	static constexpr uint32_t GetTrainBatchSize() { return kTrainBatchSize; }
	static constexpr uint32_t GetWeightCount() { return kNNWeighCount; }
	static constexpr float GetDefaultTrainProbability() { return kDefaultTrainProbability; }

	// ---- (re)initialisation: ResetMLPBuffers (src/VkNRCState.cpp:46-88)
	int ResetMLPBuffers(uint64_t seed);
	int SetWeights(const float *fp32_weights);
	int Download(uint16_t *weights, uint16_t *use_weights, void *entries, void *state, float *gradients, cudaStream_t stream);

	// ---- hot path
	int Infer(InferParams p, const void *encoded_inputs, const __half *weights, cudaStream_t stream);
	// The reference harness' whole sequence for one batch of pre-encoded queries - inputs host -> device, launch, outputs
	// device -> host (test/main.cpp:103-128) - as ONE call on host buffers: the queries are cut in chunks and the three
	// stages run on three streams (both copy engines + the SMs busy at once), ordered after / before `stream`.
	int InferEncodedHost(const void *h_inputs, void *h_outputs, uint64_t n, int clamp_output, const __half *weights, cudaStream_t stream);
	// ... and for host arrays of the reference's query records (PackedNRCInput at `input_offset` of every `stride_bytes`)
	int InferPackedHost(const void *h_records, uint32_t stride_bytes, uint32_t input_offset, void *h_outputs, uint64_t n, const NrcScene &scene,
	                    const __half *weights, cudaStream_t stream);
	// One cooperative launch of nrc_train_kernel: tp.batch[0..num_batches) (inputs / targets / counts / loss), tp.adam_mode,
	// tp.accumulate / limit / batch_cap and tp.gradients are the caller's; partials, barrier words and the optimizer
	// buffers are filled in here. `weights` is the fp16 buffer the forward / backward passes read.
	int Train(TrainParams tp, const void *encoded_inputs, const __half *weights, cudaStream_t stream);
	// ---- multi-GPU: one process per GPU; after CommConnect every Train launch all-reduces the reduced gradient with
	// the peers inside the kernel (peer-mapped inboxes over NVLink) before the replicated optimizer step.
	int CommInit(uint32_t rank, uint32_t world, cudaIpcMemHandle_t *out_handle); // allocates the local inbox
	int CommConnect(const cudaIpcMemHandle_t *all_handles);                      // world handles in rank order
	int CommAttach(uint32_t rank, uint32_t world, void *const *inboxes, void *multicast); // caller-owned buffers (+ optional NVLS mapping)
	int CommShutdown();
	int CommStatus(cudaStream_t stream);             // NRC_ERR_PEER_TIMEOUT after an exchange gave up on a peer
	void CommSetTimeout(uint32_t polls) { m_comm_spin_limit = polls ? polls : 1u; }
	uint32_t comm_world() const { return m_comm_connected ? m_comm_world : 1; }
	int AdamStep(bool write_use_weights, cudaStream_t stream);
	int SgdStep(float lr, float batch, cudaStream_t stream);
	void SetPredictionCapture(float *d) { m_prediction_capture = d; }
	float *GetPredictionCapture() const { return m_prediction_capture; }

	int device() const { return m_device; }
	int sm_count() const { return m_sms; }

private:
	int fail(int code, const std::string &what);
	int upload_initial(const float *fp32_weights);
	template <class LaunchChunk>
	int host_pipeline(const void *h_in, uint32_t in_bytes, void *h_out, uint32_t out_bytes, uint64_t n, cudaStream_t stream, LaunchChunk launch_chunk);
	int cached_map(CUtensorMap *tm, const void *base, uint64_t rows, uint32_t box_rows, std::string *err);
	struct MapEntry {
		const void *base;
		uint64_t rows, stamp;
		uint32_t box;
		CUtensorMap map;
	};
	static constexpr size_t kMapCacheEntries = 32;
	std::vector<MapEntry> m_maps;
	uint64_t m_map_clock{0};

	int m_device{0}, m_sms{0};
	Extent2D m_extent{};
	bool m_ok{false};
	int m_error_code{0};
	std::string m_error;

	__half *m_weights{nullptr}, *m_use_weights{nullptr};
	NrcOptimizerState *m_optimizer_state{nullptr};
	NrcOptimizerEntry *m_optimizer_entries{nullptr};
	float *m_gradients{nullptr}, *m_partials{nullptr};
	uint32_t *m_sync_words{nullptr}; // [0] optimizer "last CTA" counter, [2] grid-barrier arrival counter (monotonic), [3] its
	                                 // base for the next launch, [4] multi-GPU exchange epochs used so far, [5] exchange error
	                                 // flag - all device-resident, so a captured CUDA graph of the training calls replays correctly
	float *m_prediction_capture{nullptr};

	// host-buffer path: device staging (grown on demand), copy-in / copy-out streams, per-chunk events
	static constexpr int kHostChunks = 8;
	void *m_stage_in{nullptr}, *m_stage_out{nullptr};
	uint64_t m_stage_in_bytes{0}, m_stage_out_bytes{0};
	cudaStream_t m_stream_in{nullptr}, m_stream_out{nullptr};
	cudaEvent_t m_ev_start{nullptr}, m_ev_in[kHostChunks]{}, m_ev_done[kHostChunks]{}, m_ev_out{nullptr};

	uint64_t *m_comm_local{nullptr}, *m_comm_multicast{nullptr};
	uint64_t *m_comm_inbox[NRC_MAX_RANKS]{};
	uint32_t m_comm_rank{0}, m_comm_world{1};
	uint32_t m_comm_spin_limit{1u << 24}; // ~ a second of polling per word before a peer is declared missing
	bool m_comm_connected{false}, m_comm_owned{false};

	uint32_t m_seed{0};
	std::mt19937 m_rng;
	bool m_use_ema_weights{false};
	float m_train_probability{kDefaultTrainProbability};
};

// shared by the handle-less test-harness entry points
int make_weight_tensor_map(CUtensorMap *tm, const void *d_weights, std::string *err);
int make_input_tensor_map(CUtensorMap *tm, const void *d_inputs, uint64_t rows, std::string *err);

} // namespace nrc
