// nrc_b200.hpp -- header-only C++ face of the C ABI (nrc_b200.h), shaped like the reference's host object so that the
// code that talks to `VkNRCState` reads the same against this class:
//   VkNRCState                     src/VkNRCState.hpp:17-90   (ctor, Get*Buffer, ResetMLPBuffers, Set/IsUseEMAWeights,
//                                                              Set/GetTrainProbability, NextFrame / GetSeed, size statics)
//   NNInference / NNTrain passes   src/rg/NNInference.cpp:11-71, src/rg/NNTrain.hpp:95-127  -> Infer() / TrainFrame()
//   the frame order                src/rg/NRCRenderGraph.cpp:46-80                           -> Frame()
// Everything here forwards to the extern "C" entry points; nothing but <nrc_b200.h> and the CUDA runtime's stream type
// is needed to compile it. Failures throw nrc::Error (the C ABI itself never throws).
#ifndef NRC_B200_HPP
#define NRC_B200_HPP
#include "nrc_b200.h"

#include <cstdint>
#include <stdexcept>
#include <string>

namespace nrc {

struct Error : std::runtime_error {
	int code;
	Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};
inline void Check(int rc) {
	if (rc != NRC_OK)
		throw Error(rc, nrc_last_error());
}

struct Extent {
	uint32_t width, height;
};

// the per-frame resources the reference's render graph owns (src/rg/NRCRenderGraph.cpp:139-175), as device pointers
struct FrameBuffers {
	const void *eval_records = nullptr; // NRCEvalRecord[GetEvalRecordCount]           (NNInference binding 8)
	const uint32_t *eval_count = nullptr; // device-resident u32                         (binding 9)
	uint64_t max_eval_count = 0;
	void *bias_factor_r = nullptr;    // rgba32f screen image, read-modify-write        (binding 11)
	const void *factor_gb = nullptr;  // rg32f screen image                             (binding 12)
	uint32_t image_pitch = 0;         // pixels per image row
	void *train_records[4] = {};      // NRCTrainRecord[16384] x 4                      (binding 13 / NNTrain binding 8)
	uint32_t *train_counts[4] = {};   // device-resident u32 x 4, clamped in place      (NNTrain binding 9)
};

class State {
public:
	State(int device, Extent extent, uint64_t seed) { // VkNRCState(queue, extent); the seed is explicit
		nrc_config_t cfg{extent.width, extent.height, seed};
		Check(nrc_create(&cfg, device, &m_handle));
	}
	~State() { nrc_destroy(m_handle); }
	State(const State &) = delete;
	State &operator=(const State &) = delete;
	State(State &&o) noexcept : m_handle(o.m_handle) { o.m_handle = nullptr; }

	nrc_handle_t Handle() const { return m_handle; }

	// ---- src/VkNRCState.hpp:43-46 (device pointers instead of myvk::Ptr<myvk::Buffer>)
	void *GetWeightBuffer() const { return nrc_get_weight_buffer(m_handle); }
	void *GetUseWeightBuffer() const { return nrc_get_use_weight_buffer(m_handle); }
	void *GetOptimizerEntryBuffer() const { return nrc_get_optimizer_entry_buffer(m_handle); }
	void *GetOptimizerStateBuffer() const { return nrc_get_optimizer_state_buffer(m_handle); }
	void *GetGradientBuffer() const { return nrc_get_gradient_buffer(m_handle); }

	// ---- src/VkNRCState.hpp:48-81
	void ResetMLPBuffers(uint64_t seed) { Check(nrc_reset_mlp_buffers(m_handle, seed)); }
	void SetWeights(const float *fp32_weights) { Check(nrc_set_weights(m_handle, fp32_weights)); }
	bool IsUseEMAWeights() const { return nrc_is_use_ema_weights(m_handle) != 0; }
	void SetUseEMAWeights(bool v) { nrc_set_use_ema_weights(m_handle, v ? 1 : 0); }
	float GetTrainProbability() const { return nrc_get_train_probability(m_handle); }
	void SetTrainProbability(float p) { nrc_set_train_probability(m_handle, p); }
	uint32_t NextFrame() { return nrc_next_frame(m_handle); }
	uint32_t GetSeed() const { return nrc_get_seed(m_handle); }

	// ---- src/VkNRCState.hpp:83-89
	static uint64_t GetEvalRecordBufferSize(Extent e) { return nrc_get_eval_record_buffer_size(e.width, e.height); }
	static uint64_t GetBatchTrainRecordBufferSize() { return nrc_get_batch_train_record_buffer_size(); }
	static uint32_t GetTrainBatchCount() { return nrc_get_train_batch_count(); }
	static uint32_t GetTrainBatchSize() { return nrc_get_train_batch_size(); }
	static uint32_t GetWeightCount() { return nrc_get_weight_count(); }
	static float GetDefaultTrainProbability() { return nrc_get_default_train_probability(); }

	// once per scene (the reference never edits geometry after load, src/VkScene.cpp:64-133): `table` is caller-owned
	// device memory of nrc_scene_prim_table_bytes(primitives)
	static void PrepareScene(NrcScene &scene, uint32_t primitives, void *table, void *stream) {
		Check(nrc_scene_build_prim_table(&scene, primitives, table, stream));
		scene.prim_table = table;
	}

	// NRCInputEncode as a stage of its own (NRCRecord.glsl:77-95, behind UnpackNRCInput :98-125 for packed records):
	// -> [n][64] fp16 rows, the input layout of test/evaluate_NV.comp / nrc_infer_encoded / nrc_gradient_encoded
	static void EncodeInputs(const void *inputs14, uint32_t stride_bytes, uint64_t n, void *encoded, void *stream) {
		Check(nrc_encode_inputs(inputs14, stride_bytes, n, encoded, stream));
	}
	static void EncodePackedInputs(const void *packed_inputs, uint32_t stride_bytes, uint64_t n, const NrcScene &scene, void *encoded, void *stream) {
		Check(nrc_encode_packed_inputs(packed_inputs, stride_bytes, n, &scene, encoded, stream));
	}

	// the NNInference pass (nrc_inference.comp): queries -> use_weights MLP -> screen composite / train-record feedback
	void Infer(const FrameBuffers &f, const NrcScene &scene, void *stream) {
		Check(nrc_infer(m_handle, f.eval_records, f.eval_count, f.max_eval_count, &scene, f.bias_factor_r, f.factor_gb, f.image_pitch,
		                const_cast<void *const *>(f.train_records), stream));
	}
	// the four NNTrain pass groups of a frame (clear -> prepare -> gradient -> optimize, use_weights from the last), one launch
	void TrainFrame(const FrameBuffers &f, const NrcScene &scene, void *stream) {
		Check(nrc_train_frame(m_handle, const_cast<void *const *>(f.train_records), const_cast<uint32_t *const *>(f.train_counts),
		                      nrc_get_train_batch_size(), &scene, stream));
	}
	// one frame in the order of src/rg/NRCRenderGraph.cpp:46-80: inference with last frame's use_weights (including the
	// write-back into the train targets), then training - two kernel launches
	void Frame(const FrameBuffers &f, const NrcScene &scene, void *stream) {
		Check(nrc_frame(m_handle, f.eval_records, f.eval_count, f.max_eval_count, &scene, f.bias_factor_r, f.factor_gb, f.image_pitch,
		                const_cast<void *const *>(f.train_records), const_cast<uint32_t *const *>(f.train_counts), stream));
	}
	// NRCRenderGraph::PreExecute (:108-112): zero the frame's counters before the record producer runs
	void FrameBegin(const FrameBuffers &f, void *stream) {
		Check(nrc_frame_begin(m_handle, const_cast<uint32_t *>(f.eval_count), const_cast<uint32_t *const *>(f.train_counts), stream));
	}

	void Download(uint16_t *weights, uint16_t *use_weights, NrcOptimizerEntry *entries, NrcOptimizerState *state, float *gradients, void *stream) {
		Check(nrc_download(m_handle, weights, use_weights, entries, state, gradients, stream));
	}

private:
	nrc_handle_t m_handle = nullptr;
};

} // namespace nrc
#endif // NRC_B200_HPP
