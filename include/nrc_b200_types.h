/*
 * nrc_b200_types.h -- plain-C buffer layouts shared by the C ABI (nrc_b200.h) and the CUDA sources. Every struct is
 * bit-identical to the reference's (paths relative to the VkNRC tree):
 *   records      shader/src/NRCRecord.glsl:6-45, src/VkNRCState.cpp:10-32
 *   scene        shader/src/Scene.glsl:8-71, src/VkScene.hpp:22-30, src/VkScene.cpp:63-133
 */
#ifndef NRC_B200_TYPES_H
#define NRC_B200_TYPES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
typedef struct NrcPackedInput { /* NRCRecord.glsl:6-10 */
	uint32_t primitive_id, flip_bit_instance_id, barycentric_2x16U, scattered_dir_2x16U;
} NrcPackedInput;
typedef struct NrcEvalRecord { /* NRCRecord.glsl:12-18, 20 B */
	uint32_t dst;
	NrcPackedInput packed_input;
} NrcEvalRecord;
typedef struct NrcTrainRecord { /* NRCRecord.glsl:35-38, 40 B */
	float bias_r, bias_g, bias_b, factor_r, factor_g, factor_b;
	NrcPackedInput packed_input;
} NrcTrainRecord;
typedef struct NrcUnpackedInput { /* NRCRecord.glsl:40-45 flattened, 56 B */
	float position[3], scattered_dir[2], normal[2], roughness, diffuse[3], specular[3];
} NrcUnpackedInput;
/* Scene buffers the unpack step gathers from (shader/src/Scene.glsl:8-71; host mirrors src/VkScene.hpp:22-30, upload
 * src/VkScene.cpp:64-133). Plain device pointers; layouts are the reference's std430 ones. */
typedef struct NrcMaterial { /* Scene.glsl:12-20: 60 B of fields, 64 B array stride */
	float diffuse[3];
	uint32_t diffuse_texture_id; /* 0xFFFFFFFF = none (Scene.glsl:60) */
	float specular[3];
	uint32_t specular_texture_id;
	float emission[3];
	uint32_t emission_texture_id;
	float metallic, roughness, ior;
	uint32_t _pad;
} NrcMaterial;
typedef struct NrcTexture { /* R8G8B8A8_SRGB, one mip level (src/VkScene.cpp:193), sampled LINEAR + REPEAT (NRCRenderGraph.cpp:132) */
	const void *texels_rgba8_srgb;
	uint32_t width, height;
} NrcTexture;
typedef struct NrcScene {
	const float *vertices;            /* Vertex{x,y,z}, 12 B each (Scene.glsl:8-10) */
	const uint32_t *vertex_indices;   /* 3 per primitive */
	const float *texcoords;           /* vec2 */
	const uint32_t *texcoord_indices; /* 3 per primitive */
	const NrcMaterial *materials;
	const uint32_t *material_ids;     /* per primitive */
	const float *transforms;          /* mat3x4 per instance = 3 vec4 columns {r0.xyz t.x | r1.xyz t.y | r2.xyz t.z} (VkScene.cpp:63-71) */
	const NrcTexture *textures;       /* device array of descriptors */
	uint32_t texture_count;
	/* Optional (NULL = gather through the index buffers as the reference does): NrcPrimRow per primitive, built once per
	 * scene by nrc_scene_build_prim_table - one 64-byte line instead of 7 index + 15 attribute loads on two dependent
	 * levels. The values are copies, so results are bit-identical either way. */
	const void *prim_table;
} NrcScene;
typedef struct NrcPrimRow { /* 64 B, 64-byte aligned */
	float v[3][3];        /* object-space vertices (GetSceneVertex, Scene.glsl:50-52) */
	float tc[3][2];       /* texture coordinates (GetSceneTexcoord, Scene.glsl:53-56) */
	uint32_t material_id;
} NrcPrimRow;
typedef struct NrcOptimizerState { /* src/VkNRCState.cpp:25-28, 20 B */
	uint32_t t;
	float beta1_t, beta2_t, alpha_t, alpha_t_1;
} NrcOptimizerState;
typedef struct NrcOptimizerEntry { /* src/VkNRCState.cpp:29-31, 16 B */
	float m, v, weight, ema_weight;
} NrcOptimizerEntry;
#ifdef __cplusplus
}
#endif

#endif /* NRC_B200_TYPES_H */
