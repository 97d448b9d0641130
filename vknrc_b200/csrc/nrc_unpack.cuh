// nrc_unpack.cuh -- UnpackNRCInput (shader/src/NRCRecord.glsl:98-125): a 16-byte PackedNRCInput -> the 14 floats the
// encoder eats, gathered per thread from the scene buffers of shader/src/Scene.glsl:8-71 (vertices through the index and
// per-instance transform buffers, texture coordinates, the per-primitive material, up to two texture fetches).
// Textures are read with the semantics of the reference's sampler (R8G8B8A8_SRGB, one mip level, VK_FILTER_LINEAR,
// ADDRESS_MODE_REPEAT - src/VkScene.cpp:193, src/rg/NRCRenderGraph.cpp:132) in software: sRGB -> linear per texel, then a
// bilinear blend with fp32 weights, so the result does not depend on the texture unit's 8-bit weight quantisation.
#pragma once
#include "nrc_config.h"

namespace nrc {

__device__ __forceinline__ float srgb_to_linear(uint32_t c8) {
	const float x = (float)c8 * (1.0f / 255.0f);
	return x <= 0.04045f ? x * (1.0f / 12.92f) : powf((x + 0.055f) * (1.0f / 1.055f), 2.4f);
}

__device__ __forceinline__ void sample_texture_srgb_repeat(const NrcTexture *tex, float u, float v, float rgb[3]) {
	const uint32_t *p = (const uint32_t *)tex->texels_rgba8_srgb;
	const int w = (int)tex->width, h = (int)tex->height;
	const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
	const float fx = floorf(x), fy = floorf(y), tx = x - fx, ty = y - fy;
	int x0 = (int)fx % w, y0 = (int)fy % h;
	x0 += x0 < 0 ? w : 0, y0 += y0 < 0 ? h : 0;
	const int x1 = x0 + 1 == w ? 0 : x0 + 1, y1 = y0 + 1 == h ? 0 : y0 + 1;
	const uint32_t c00 = __ldg(p + y0 * w + x0), c10 = __ldg(p + y0 * w + x1), c01 = __ldg(p + y1 * w + x0), c11 = __ldg(p + y1 * w + x1);
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const float a = srgb_to_linear((c00 >> (8 * c)) & 255u), b = srgb_to_linear((c10 >> (8 * c)) & 255u);
		const float d = srgb_to_linear((c01 >> (8 * c)) & 255u), e = srgb_to_linear((c11 >> (8 * c)) & 255u);
		rgb[c] = (a * (1.0f - tx) + b * tx) * (1.0f - ty) + (d * (1.0f - tx) + e * tx) * ty;
	}
}

__device__ __forceinline__ void unpack_nrc_input(const NrcScene &sc, const uint32_t pk[4], float out[14]) {
	const uint32_t prim = pk[0], instance = pk[1] & 0x7FFFFFFFu;
	const bool flip = (pk[1] >> 31) != 0u;
	const float4 *m = (const float4 *)sc.transforms + 3 * (size_t)instance; // vec4(v, 1) * mat3x4 = a dot product per column
	const float4 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
	float v[3][3], tc[3][2];
#pragma unroll
	for (int k = 0; k < 3; ++k) { // GetSceneVertex / GetSceneTexcoord (Scene.glsl:50-56)
		const float *p = sc.vertices + 3 * (size_t)__ldg(sc.vertex_indices + 3 * (size_t)prim + k);
		const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
		v[k][0] = x * m0.x + y * m0.y + z * m0.z + m0.w;
		v[k][1] = x * m1.x + y * m1.y + z * m1.z + m1.w;
		v[k][2] = x * m2.x + y * m2.y + z * m2.z + m2.w;
		const float2 t = __ldg((const float2 *)sc.texcoords + __ldg(sc.texcoord_indices + 3 * (size_t)prim + k));
		tc[k][0] = t.x, tc[k][1] = t.y;
	}
	const float e1x = v[1][0] - v[0][0], e1y = v[1][1] - v[0][1], e1z = v[1][2] - v[0][2];
	const float e2x = v[2][0] - v[0][0], e2y = v[2][1] - v[0][1], e2z = v[2][2] - v[0][2];
	float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
	const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
	nx *= inv, ny *= inv, nz *= inv;
	if (flip)
		nx = -nx, ny = -ny, nz = -nz;
	const float by = (float)(pk[2] & 0xFFFFu) / 65535.0f, bz = (float)(pk[2] >> 16) / 65535.0f, bx = 1.0f - by - bz; // :111-112
#pragma unroll
	for (int j = 0; j < 3; ++j)
		out[j] = v[0][j] * bx + v[1][j] * by + v[2][j] * bz;
	out[3] = (float)(pk[3] & 0xFFFFu) / 65535.0f, out[4] = (float)(pk[3] >> 16) / 65535.0f; // :118
	const float kPi = 3.14159265358979323846f;
	out[5] = (nx == 0.0f && ny == 0.0f) ? 0.5f : 0.5f + atan2f(ny, nx) / (2.0f * kPi); // NRCSphEncode, :47-49
	out[6] = acosf(fminf(fmaxf(nz, -1.0f), 1.0f)) / kPi;
	const NrcMaterial *mat = sc.materials + __ldg(sc.material_ids + prim);
	const float4 md = __ldg((const float4 *)mat), ms = __ldg((const float4 *)mat + 1);
	out[7] = __ldg(&mat->roughness);
	const float u = tc[0][0] * bx + tc[1][0] * by + tc[2][0] * bz, w = tc[0][1] * bx + tc[1][1] * by + tc[2][1] * bz;
	const uint32_t dtex = __float_as_uint(md.w), stex = __float_as_uint(ms.w);
	out[8] = md.x, out[9] = md.y, out[10] = md.z, out[11] = ms.x, out[12] = ms.y, out[13] = ms.z;
	if (dtex != 0xFFFFFFFFu) // GetSceneDiffuse / GetSceneSpecular (Scene.glsl:59-64)
		sample_texture_srgb_repeat(sc.textures + dtex, u, w, out + 8);
	if (stex != 0xFFFFFFFFu)
		sample_texture_srgb_repeat(sc.textures + stex, u, w, out + 11);
}

__device__ __forceinline__ void load_packed_input(const void *base, uint64_t index, uint32_t stride_bytes, uint32_t pk[4]) {
	const uint32_t *p = (const uint32_t *)((const uint8_t *)base + index * stride_bytes); // 4-byte aligned only (20 / 40 B records)
	pk[0] = __ldg(p), pk[1] = __ldg(p + 1), pk[2] = __ldg(p + 2), pk[3] = __ldg(p + 3);
}

} // namespace nrc
