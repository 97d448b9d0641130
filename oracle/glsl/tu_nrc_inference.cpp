// TEST INFRASTRUCTURE: shader/src/nrc_inference.comp (+ NRCRecord.glsl, Scene.glsl, NN_nv.glsl) of the reference, compiled as C++.
#include "glsl_api.h"
#include "glsl_runtime.hpp"
#include "glsl_shim.hpp"
#include <vector>
#define SUBGROUP_SIZE 32 // shader/CMakeLists.txt:37-43 builds the 16 and 32 variants; B200-class hardware runs 32
namespace sh_inference {
#include "shader/src/nrc_inference.comp"
#include "tu_scene_common.hpp"
} // namespace sh_inference

extern "C" {
void glsl_NRCInputEncode(const float *in, uint64_t n, uint16_t *out64) {
	for (uint64_t i = 0; i < n; ++i) {
		const float *p = in + 14 * i;
		sh_inference::UnpackedNRCInput u;
		u.position = vec3(p[0], p[1], p[2]), u.scattered_dir = vec2(p[3], p[4]), u.normal = vec2(p[5], p[6]), u.roughness = p[7];
		u.diffuse = vec3(p[8], p[9], p[10]), u.specular = vec3(p[11], p[12], p[13]);
		uvec4 o[8];
		sh_inference::NRCInputEncode(u, o);
		std::memcpy(out64 + 64 * i, o, 128);
	}
}
void glsl_UnpackNRCInput(const GlslScene *scene, const uint32_t *packed4, uint64_t n, float *out14) {
	sh_inference::SceneBinding b;
	b.bind(scene);
	for (uint64_t i = 0; i < n; ++i) {
		sh_inference::PackedNRCInput pk{packed4[4 * i], packed4[4 * i + 1], packed4[4 * i + 2], packed4[4 * i + 3]};
		const sh_inference::UnpackedNRCInput u = sh_inference::UnpackNRCInput(pk);
		float *o = out14 + 14 * i;
		o[0] = u.position.x, o[1] = u.position.y, o[2] = u.position.z, o[3] = u.scattered_dir.x, o[4] = u.scattered_dir.y;
		o[5] = u.normal.x, o[6] = u.normal.y, o[7] = u.roughness;
		o[8] = u.diffuse.x, o[9] = u.diffuse.y, o[10] = u.diffuse.z, o[11] = u.specular.x, o[12] = u.specular.y, o[13] = u.specular.z;
	}
}
uint32_t glsl_EncodeNRCEvalDstScreen(uint32_t x, uint32_t y) { return sh_inference::EncodeNRCEvalDstScreen(uvec2(x, y)); }
uint32_t glsl_EncodeNRCEvalDstTrain(uint32_t b, uint32_t l, uint32_t r) { return sh_inference::EncodeNRCEvalDstTrain(b, l, r); }
void glsl_DecodeNRCEvalDst(uint32_t e, uint32_t *type, uint32_t *a, uint32_t *b, uint32_t *c) {
	*type = sh_inference::GetNRCEvalDstType(e), *a = *b = *c = 0;
	if (*type == NRC_EVAL_DST_SCREEN) {
		const uvec2 xy = sh_inference::DecodeNRCEvalDstScreen(e);
		*a = xy.x, *b = xy.y;
	} else {
		sh_inference::DecodeNRCEvalDstTrain(e, *a, *b, *c);
	}
}
int glsl_nrc_inference(const GlslScene *scene, const void *eval_records, uint32_t eval_count, const uint16_t *weights, float *bias_factor_r,
                       const float *factor_gb, uint32_t width, uint32_t height, void *const train_records[4], int parallel) {
	static_assert(sizeof(sh_inference::NRCEvalRecord) == 20 && sizeof(sh_inference::NRCTrainRecord) == 40, "record layout");
	sh_inference::SceneBinding b;
	b.bind(scene);
	sh_inference::uEvalRecords = (sh_inference::NRCEvalRecord *)eval_records, sh_inference::uEvalCount = eval_count, sh_inference::uWeights = (uvec4 *)weights;
	sh_inference::uBias_FactorR = image2D{bias_factor_r, (int)width, (int)height, (int)width, GLSL_RGBA32F};
	sh_inference::uFactorGB = image2D{(void *)factor_gb, (int)width, (int)height, (int)width, GLSL_RG32F};
	for (int k = 0; k < NRC_TRAIN_BATCH_COUNT; ++k)
		sh_inference::uBatchTrainRecords[k].records = (sh_inference::NRCTrainRecord *)train_records[k];
	glsl_rt::dispatch((eval_count + 127) / 128, 128, SUBGROUP_SIZE, &sh_inference::main, parallel != 0); // nrc_indirect.comp:10
	return 0;
}
}
