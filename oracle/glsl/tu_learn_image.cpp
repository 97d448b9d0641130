// TEST INFRASTRUCTURE: test/mlp_learning_an_image/{gradient,optimize,inference}.comp of the reference, compiled as C++.
#include "glsl_api.h"
#include "glsl_runtime.hpp"
#include "glsl_shim.hpp"
#define SUBGROUP_SIZE 32
namespace sh_grad {
#include "test/mlp_learning_an_image/gradient.comp"
}
#undef NN_NV_GLSL
#undef NN_BACKPROPAGATION
#undef WEIGHTS_BINDING
#undef DWEIGHTS_BINDING
#undef WORKGROUP_SIZE
#undef SHARED_BUFFER_SIZE
namespace sh_inf {
#include "test/mlp_learning_an_image/inference.comp"
}
namespace sh_opt {
#include "test/mlp_learning_an_image/optimize.comp"
}
extern "C" {
int glsl_image_gradient(const uint16_t *weights, float *dweights, const void *image_rgba8, uint32_t w, uint32_t h, uint32_t seed_x, uint32_t seed_y,
                        uint32_t n, int parallel) {
	if (n % 128)
		return -1;
	sh_grad::uWeights = (uvec4 *)weights, sh_grad::uDWeights = dweights, sh_grad::uSeed = uvec2(seed_x, seed_y);
	sh_grad::uImage = sampler2D{(const uint8_t *)image_rgba8, (int)w, (int)h, /*srgb=*/false, /*repeat=*/false}; // main.cpp:43, 121-124
	glsl_rt::dispatch(n / 128, 128, SUBGROUP_SIZE, &sh_grad::main, parallel != 0);
	return 0;
}
void glsl_image_optimize(uint16_t *weights, float *fp_weights, const float *gradients) {
	sh_opt::uWeights = (float16_t *)weights, sh_opt::uFPWeights = fp_weights, sh_opt::uGradients = (float *)gradients;
	glsl_rt::dispatch(20672 / 64, 64, 32, &sh_opt::main); // main.cpp:172
}
int glsl_image_inference(const uint16_t *weights, void *out_rgba8, int parallel) {
	sh_inf::uWeights = (uvec4 *)weights;
	sh_inf::uOutput = image2D{out_rgba8, WINDOW_SIZE, WINDOW_SIZE, WINDOW_SIZE, GLSL_RGBA8};
	glsl_rt::dispatch(WINDOW_SIZE * WINDOW_SIZE / 128, 128, SUBGROUP_SIZE, &sh_inf::main, parallel != 0); // main.cpp:214
	return 0;
}
void glsl_image_uv(uint32_t seed_x, uint32_t seed_y, uint32_t gid, float *u, float *v) { // the expression of gradient.comp:47-48
	const vec2 uv = (1.0f / float(0xffffffffu)) * vec2(sh_grad::pcg2d(uvec2(seed_x + gid % 128, seed_y + gid / 128)));
	*u = uv.x, *v = uv.y;
}
void glsl_image_oneblob32(float u, float v, uint16_t *out64) { // gradient.comp:51-56
	float ob_u[32], ob_v[32];
	sh_grad::oneblob_32(u, ob_u), sh_grad::oneblob_32(v, ob_v);
	uvec4 inputs[8];
	sh_grad::pack_half_32(ob_u, 0, inputs), sh_grad::pack_half_32(ob_v, 4, inputs);
	std::memcpy(out64, inputs, 128);
}
}
