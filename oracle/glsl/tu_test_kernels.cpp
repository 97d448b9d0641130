// TEST INFRASTRUCTURE: test/evaluate_NV.comp and test/train_NV.comp of the reference (with test/NN_nv.glsl), compiled as C++.
#include "glsl_api.h"
#include "glsl_runtime.hpp"
#include "glsl_shim.hpp"
#define SUBGROUP_SIZE 32
namespace sh_test_eval {
#include "test/evaluate_NV.comp"
}
#undef NN_NV_GLSL
#undef WEIGHTS_BINDING
#undef WORKGROUP_SIZE
#undef SHARED_BUFFER_SIZE
namespace sh_test_train {
#include "test/train_NV.comp"
}
extern "C" {
int glsl_evaluate_NV(const uint16_t *weights, const uint16_t *inputs, uint64_t n, uint16_t *outputs3, int parallel) {
	if (n % 128)
		return -1; // the test kernels have no tail guard
	static_assert(sizeof(sh_test_eval::F16Vec3) == 6, "F16Vec3");
	sh_test_eval::uWeights = (uvec4 *)weights, sh_test_eval::uInputs = (uvec4 *)inputs, sh_test_eval::uOutputs = (sh_test_eval::F16Vec3 *)outputs3;
	glsl_rt::dispatch((uint32_t)(n / 128), 128, SUBGROUP_SIZE, &sh_test_eval::main, parallel != 0);
	return 0;
}
int glsl_train_NV(const uint16_t *weights, float *dweights, const uint16_t *inputs, const uint16_t *targets3, uint64_t n, int parallel) {
	if (n % 128)
		return -1;
	sh_test_train::uWeights = (uvec4 *)weights, sh_test_train::uDWeights = dweights, sh_test_train::uInputs = (uvec4 *)inputs;
	sh_test_train::uTargets = (sh_test_train::F16Vec3 *)targets3;
	glsl_rt::dispatch((uint32_t)(n / 128), 128, SUBGROUP_SIZE, &sh_test_train::main, parallel != 0);
	return 0;
}
}
