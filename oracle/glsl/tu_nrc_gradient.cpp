// TEST INFRASTRUCTURE: shader/src/nrc_gradient.comp of the reference, compiled as C++.
#include "glsl_api.h"
#include "glsl_runtime.hpp"
#include "glsl_shim.hpp"
#include <vector>
#define SUBGROUP_SIZE 32
namespace sh_gradient {
#include "shader/src/nrc_gradient.comp"
#include "tu_scene_common.hpp"
} // namespace sh_gradient
extern "C" int glsl_nrc_gradient(const GlslScene *scene, const void *train_records, uint32_t count, const uint16_t *weights, float *dweights, int parallel) {
	sh_gradient::SceneBinding b;
	b.bind(scene);
	sh_gradient::uBatchTrainRecords = (sh_gradient::NRCTrainRecord *)train_records, sh_gradient::uBatchTrainCount = count;
	sh_gradient::uWeights = (uvec4 *)weights, sh_gradient::uDWeights = dweights;
	glsl_rt::dispatch((count + 127) / 128, 128, SUBGROUP_SIZE, &sh_gradient::main, parallel != 0); // nrc_train_prepare.comp:21
	return 0;
}
