/* glsl_api.h -- TEST INFRASTRUCTURE: C entry points of oracle/_ref_glsl/libvknrc_glsl.so, the reference's own GLSL
 * shaders compiled as C++ (glsl2cpp.py + glsl_shim.hpp) and run on the CPU. Every function names the shader it runs. */
#ifndef VKNRC_GLSL_API_H
#define VKNRC_GLSL_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct GlslTexture {
	const void *texels_rgba8; /* R8G8B8A8_SRGB */
	uint32_t width, height;
} GlslTexture;
typedef struct GlslScene { /* the buffers of shader/src/Scene.glsl:24-46 in the layouts src/VkScene.cpp uploads */
	const float *vertices;
	const uint32_t *vertex_indices;
	const float *texcoords;
	const uint32_t *texcoord_indices;
	const void *materials; /* 64-byte std430 stride */
	const uint32_t *material_ids;
	const float *transforms; /* mat3x4 per instance */
	const GlslTexture *textures;
	uint32_t texture_count, material_count;
} GlslScene;
typedef struct GlslOptimizerState {
	uint32_t step;
	float beta1_t, beta2_t, alpha_t, alpha_t_1;
} GlslOptimizerState;

/* shader/src/NRCRecord.glsl: functions called directly */
void glsl_NRCInputEncode(const float *unpacked14, uint64_t n, uint16_t *out64);                 /* :78-95 */
void glsl_UnpackNRCInput(const GlslScene *scene, const uint32_t *packed4, uint64_t n, float *out14); /* :98-125 */
uint32_t glsl_EncodeNRCEvalDstScreen(uint32_t x, uint32_t y);                                    /* :19 */
uint32_t glsl_EncodeNRCEvalDstTrain(uint32_t b, uint32_t l, uint32_t r);                         /* :20-22 */
void glsl_DecodeNRCEvalDst(uint32_t e, uint32_t *type, uint32_t *a, uint32_t *b, uint32_t *c);   /* :23-33 */
/* whole shaders (main() dispatched over ceil(n / 128) workgroups of 128 invocations, subgroup size 32) */
int glsl_evaluate_NV(const uint16_t *weights, const uint16_t *inputs, uint64_t n, uint16_t *outputs3, int parallel); /* test/evaluate_NV.comp */
int glsl_train_NV(const uint16_t *weights, float *dweights, const uint16_t *inputs, const uint16_t *targets3, uint64_t n, int parallel); /* test/train_NV.comp */
int glsl_nrc_inference(const GlslScene *scene, const void *eval_records, uint32_t eval_count, const uint16_t *weights, float *bias_factor_r,
                       const float *factor_gb, uint32_t width, uint32_t height, void *const train_records[4], int parallel); /* nrc_inference.comp */
int glsl_nrc_gradient(const GlslScene *scene, const void *train_records, uint32_t count, const uint16_t *weights, float *dweights, int parallel); /* nrc_gradient.comp */
void glsl_nrc_train_prepare(uint32_t *count, uint32_t command[3], GlslOptimizerState *state);    /* nrc_train_prepare.comp */
void glsl_nrc_optimize(uint16_t *weights, uint16_t *use_weights /* NULL: the variant without WRITE_USE_WEIGHTS */, const float *gradients,
                       void *optimizer_entries, uint32_t count, const GlslOptimizerState *state, uint32_t use_ema_weights); /* nrc_optimize.comp */
/* test/mlp_learning_an_image */
int glsl_image_gradient(const uint16_t *weights, float *dweights, const void *image_rgba8, uint32_t w, uint32_t h, uint32_t seed_x, uint32_t seed_y,
                        uint32_t n, int parallel);                                               /* gradient.comp */
void glsl_image_optimize(uint16_t *weights, float *fp_weights, const float *gradients);          /* optimize.comp */
int glsl_image_inference(const uint16_t *weights, void *out_rgba8, int parallel);                /* inference.comp (640 x 640) */
void glsl_image_uv(uint32_t seed_x, uint32_t seed_y, uint32_t gid, float *u, float *v);          /* gradient.comp:47-48 */
void glsl_image_oneblob32(float u, float v, uint16_t *out64);                                    /* gradient.comp:33-44, 51-56 */
#ifdef __cplusplus
}
#endif
#endif
