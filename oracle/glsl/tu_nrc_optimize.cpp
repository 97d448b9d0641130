// TEST INFRASTRUCTURE: shader/src/nrc_optimize.comp (both variants, shader/CMakeLists.txt:44-45) and nrc_train_prepare.comp.
#include "glsl_api.h"
#include "glsl_runtime.hpp"
#include "glsl_shim.hpp"
namespace sh_plain {
#include "shader/src/nrc_optimize.comp"
}
#undef CONSTANT_GLSL
#undef LEARNING_RATE
#undef EPSILON
#undef M_PI
#define WRITE_USE_WEIGHTS
namespace sh_use {
#include "shader/src/nrc_optimize.comp"
}
#undef CONSTANT_GLSL
#undef M_PI
namespace sh_prepare {
#include "shader/src/nrc_train_prepare.comp"
}
extern "C" {
void glsl_nrc_train_prepare(uint32_t *count, uint32_t command[3], GlslOptimizerState *st) {
	sh_prepare::uCount = *count;
	sh_prepare::uStep = st->step, sh_prepare::uBeta1_T = st->beta1_t, sh_prepare::uBeta2_T = st->beta2_t;
	sh_prepare::uAlpha_T = st->alpha_t, sh_prepare::uAlpha_T_1 = st->alpha_t_1;
	glsl_rt::dispatch(1, 1, 1, &sh_prepare::main);
	*count = sh_prepare::uCount;
	command[0] = sh_prepare::uCommand.x, command[1] = sh_prepare::uCommand.y, command[2] = sh_prepare::uCommand.z;
	st->step = sh_prepare::uStep, st->beta1_t = sh_prepare::uBeta1_T, st->beta2_t = sh_prepare::uBeta2_T;
	st->alpha_t = sh_prepare::uAlpha_T, st->alpha_t_1 = sh_prepare::uAlpha_T_1;
}
void glsl_nrc_optimize(uint16_t *weights, uint16_t *use_weights, const float *gradients, void *entries, uint32_t count, const GlslOptimizerState *st,
                       uint32_t use_ema) {
	static_assert(sizeof(sh_plain::OptimizerEntry) == 16, "entry layout");
	const uint32_t groups = 20672 / 64; // src/rg/NNTrain.cpp:132
	if (use_weights) {
		sh_use::uWeights = (float16_t *)weights, sh_use::uUseWeights = (float16_t *)use_weights, sh_use::uGradients = (float *)gradients;
		sh_use::uOptimizerEntries = (sh_use::OptimizerEntry *)entries, sh_use::uBatchTrainCount = count, sh_use::uUseEMAWeights = use_ema;
		sh_use::uStep = st->step, sh_use::uBeta1_T = st->beta1_t, sh_use::uBeta2_T = st->beta2_t, sh_use::uAlpha_T = st->alpha_t, sh_use::uAlpha_T_1 = st->alpha_t_1;
		glsl_rt::dispatch(groups, 64, 32, &sh_use::main);
	} else {
		sh_plain::uWeights = (float16_t *)weights, sh_plain::uGradients = (float *)gradients;
		sh_plain::uOptimizerEntries = (sh_plain::OptimizerEntry *)entries, sh_plain::uBatchTrainCount = count, sh_plain::uUseEMAWeights = use_ema;
		sh_plain::uStep = st->step, sh_plain::uBeta1_T = st->beta1_t, sh_plain::uBeta2_T = st->beta2_t, sh_plain::uAlpha_T = st->alpha_t, sh_plain::uAlpha_T_1 = st->alpha_t_1;
		glsl_rt::dispatch(groups, 64, 32, &sh_plain::main);
	}
}
}
