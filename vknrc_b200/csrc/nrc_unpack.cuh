// nrc_unpack.cuh -- UnpackNRCInput (shader/src/NRCRecord.glsl:98-125): a 16-byte PackedNRCInput -> the 14 floats the
// encoder eats, gathered per thread from the scene buffers of shader/src/Scene.glsl:8-71 (vertices through the index and
// per-instance transform buffers, texture coordinates, the per-primitive material, up to two texture fetches).
// Textures are read with the semantics of the reference's sampler (R8G8B8A8_SRGB, one mip level, VK_FILTER_LINEAR,
// ADDRESS_MODE_REPEAT - src/VkScene.cpp:193, src/rg/NRCRenderGraph.cpp:132) in software: sRGB -> linear per texel, then a
// bilinear blend with fp32 weights, so the result does not depend on the texture unit's 8-bit weight quantisation.
#pragma once
#include "nrc_config.h"

namespace nrc {

// One 256-bit read-only load (LDG.E.256 on sm_100): a 64-byte primitive row is two of them instead of four 128-bit loads. The
// gather is bound by the number of divergent load instructions (every lane a line of its own: one L1 tag look-up per lane
// and instruction), not by bytes: nrc_infer 131.1 -> 126.5 us at 1080p with the primitive row and the material read this way.
__device__ __forceinline__ void ldg256(const void *p32_byte_aligned, float4 &a, float4 &b) {
	asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	             : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
	             : "l"(p32_byte_aligned));
}

// sRGB -> linear for the 256 texel values: x <= 0.04045 ? x / 12.92 : ((x + 0.055) / 1.055)^2.4 with x = c / 255, evaluated
// in double precision and rounded to fp32 (tools/gen_srgb_table.py). A table, as in the texture unit - powf per channel
// per texel (24 per textured record) costs more than the whole MLP.
__device__ const float kSrgbToLinear[256] = {
	0.0f, 0.000303526991f, 0.000607053982f, 0.000910580973f, 0.00121410796f, 0.00151763496f, 0.00182116195f, 0.00212468882f,
	0.00242821593f, 0.0027317428f, 0.00303526991f, 0.00334653584f, 0.00367650739f, 0.00402471703f, 0.00439144205f, 0.00477695325f,
	0.00518151652f, 0.00560539169f, 0.00604883302f, 0.00651209056f, 0.00699541019f, 0.00749903219f, 0.00802319311f, 0.00856812578f,
	0.00913405884f, 0.00972121768f, 0.010329823f, 0.0109600937f, 0.0116122449f, 0.012286488f, 0.0129830325f, 0.0137020834f,
	0.0144438436f, 0.0152085144f, 0.0159962941f, 0.0168073755f, 0.0176419541f, 0.01850022f, 0.0193823613f, 0.0202885624f,
	0.0212190095f, 0.0221738853f, 0.0231533665f, 0.0241576321f, 0.0251868591f, 0.0262412224f, 0.0273208916f, 0.02842604f,
	0.0295568351f, 0.0307134446f, 0.0318960324f, 0.0331047662f, 0.0343398079f, 0.0356013142f, 0.0368894488f, 0.0382043719f,
	0.0395462364f, 0.0409151986f, 0.0423114114f, 0.043735031f, 0.045186203f, 0.0466650873f, 0.0481718257f, 0.0497065671f,
	0.0512694567f, 0.0528606474f, 0.054480277f, 0.0561284907f, 0.0578054301f, 0.0595112368f, 0.0612460524f, 0.0630100146f,
	0.064803265f, 0.0666259378f, 0.0684781671f, 0.0703600943f, 0.0722718537f, 0.0742135718f, 0.0761853829f, 0.078187421f,
	0.0802198201f, 0.0822827071f, 0.0843762085f, 0.0865004584f, 0.0886555836f, 0.0908417106f, 0.0930589661f, 0.0953074694f,
	0.097587347f, 0.0998987257f, 0.102241732f, 0.104616486f, 0.107023105f, 0.10946171f, 0.111932427f, 0.114435375f,
	0.116970666f, 0.119538426f, 0.122138776f, 0.124771819f, 0.127437681f, 0.130136475f, 0.13286832f, 0.135633335f,
	0.138431609f, 0.141263291f, 0.144128472f, 0.147027269f, 0.149959788f, 0.152926147f, 0.155926466f, 0.158960834f,
	0.162029371f, 0.165132195f, 0.168269396f, 0.171441108f, 0.174647406f, 0.177888423f, 0.18116425f, 0.18447499f,
	0.187820777f, 0.191201687f, 0.194617838f, 0.198069319f, 0.20155625f, 0.205078736f, 0.208636865f, 0.212230757f,
	0.215860501f, 0.219526201f, 0.223227963f, 0.226965874f, 0.230740055f, 0.23455058f, 0.238397568f, 0.242281124f,
	0.246201321f, 0.25015828f, 0.254152089f, 0.258182853f, 0.262250662f, 0.266355604f, 0.270497799f, 0.274677306f,
	0.278894275f, 0.283148736f, 0.287440836f, 0.291770637f, 0.296138257f, 0.300543785f, 0.304987311f, 0.309468925f,
	0.313988715f, 0.318546772f, 0.323143214f, 0.327778101f, 0.332451522f, 0.337163627f, 0.341914415f, 0.346704066f,
	0.351532608f, 0.356400132f, 0.361306787f, 0.366252601f, 0.371237695f, 0.376262128f, 0.38132602f, 0.386429429f,
	0.391572475f, 0.396755219f, 0.401977777f, 0.407240212f, 0.412542611f, 0.417885065f, 0.423267663f, 0.428690493f,
	0.434153646f, 0.439657182f, 0.445201188f, 0.450785786f, 0.456411034f, 0.462076992f, 0.467783809f, 0.473531485f,
	0.479320168f, 0.48514995f, 0.491020858f, 0.496932983f, 0.502886474f, 0.50888133f, 0.514917672f, 0.520995557f,
	0.527115107f, 0.533276379f, 0.539479494f, 0.545724452f, 0.55201143f, 0.558340371f, 0.564711511f, 0.571124852f,
	0.577580452f, 0.584078431f, 0.590618849f, 0.597201765f, 0.603827357f, 0.610495567f, 0.617206573f, 0.623960376f,
	0.630757153f, 0.637596846f, 0.644479692f, 0.651405632f, 0.658374846f, 0.665387273f, 0.672443151f, 0.679542482f,
	0.686685324f, 0.693871737f, 0.701101899f, 0.708375752f, 0.715693474f, 0.723055124f, 0.730460763f, 0.73791039f,
	0.745404184f, 0.752942204f, 0.760524511f, 0.768151164f, 0.775822222f, 0.783537805f, 0.791297913f, 0.799102724f,
	0.806952238f, 0.814846575f, 0.822785735f, 0.830769897f, 0.838799f, 0.846873224f, 0.854992628f, 0.863157213f,
	0.871367097f, 0.8796224f, 0.887923121f, 0.896269381f, 0.904661179f, 0.913098633f, 0.921581864f, 0.930110872f,
	0.938685715f, 0.947306514f, 0.955973327f, 0.964686275f, 0.973445296f, 0.982250571f, 0.991102099f, 1.0f,
};
// `lut` = the table above or a copy of it in shared memory (the fused kernels stage it once per CTA: 24 look-ups per
// textured record are 24 scattered L1 requests from global memory, but cheap LDS from shared memory)
// floor-modulo of an integer-valued float by the texture size, in floating point (a runtime integer modulo costs ~25
// instructions): the quotient estimate can be off by one only when `f` is a multiple of `size`, which the two fix-ups catch
__device__ __forceinline__ int wrap_repeat(float f, float size) {
	float r = fmaf(-size, floorf(f * __frcp_rn(size)), f);
	r = r >= size ? r - size : r;
	r = r < 0.0f ? r + size : r;
	return (int)r;
}
__device__ __forceinline__ void sample_texture_srgb_repeat(const NrcTexture *tex, float u, float v, float rgb[3], const float *lut) {
	const uint32_t *p = (const uint32_t *)tex->texels_rgba8_srgb;
	const int w = (int)tex->width, h = (int)tex->height;
	const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
	const float fx = floorf(x), fy = floorf(y), tx = x - fx, ty = y - fy;
	const int x0 = wrap_repeat(fx, (float)w), y0 = wrap_repeat(fy, (float)h);
	const int x1 = x0 + 1 == w ? 0 : x0 + 1, y1 = y0 + 1 == h ? 0 : y0 + 1;
	const uint32_t c00 = __ldg(p + y0 * w + x0), c10 = __ldg(p + y0 * w + x1), c01 = __ldg(p + y1 * w + x0), c11 = __ldg(p + y1 * w + x1);
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const float a = lut[(c00 >> (8 * c)) & 255u], b = lut[(c10 >> (8 * c)) & 255u];
		const float d = lut[(c01 >> (8 * c)) & 255u], e = lut[(c11 >> (8 * c)) & 255u];
		rgb[c] = (a * (1.0f - tx) + b * tx) * (1.0f - ty) + (d * (1.0f - tx) + e * tx) * ty;
	}
}

// `textures` = sc.textures or a copy of the descriptor table in shared memory (one dependent load level less)
__device__ __forceinline__ void unpack_nrc_input(const NrcScene &sc, const uint32_t pk[4], float out[14], const float *lut = kSrgbToLinear,
                                                 const NrcTexture *textures = nullptr) {
	if (!textures)
		textures = sc.textures;
	const uint32_t prim = pk[0], instance = pk[1] & 0x7FFFFFFFu;
	const bool flip = (pk[1] >> 31) != 0u;
	const float4 *m = (const float4 *)sc.transforms + 3 * (size_t)instance; // vec4(v, 1) * mat3x4 = a dot product per column
	const float4 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
	float o[3][3], v[3][3], tc[3][2]; // object-space vertices, world-space vertices, texture coordinates
	uint32_t material_id;
	if (sc.prim_table) { // one 64-byte row per primitive (nrc_scene_build_prim_table): four 16-byte loads, one level
		const float4 *row = (const float4 *)sc.prim_table + 4 * (size_t)prim;
		float4 r0, r1, r2, r3; // (the table is 64-byte aligned: nrc_scene_build_prim_table)
		ldg256(row, r0, r1), ldg256(row + 2, r2, r3);
		o[0][0] = r0.x, o[0][1] = r0.y, o[0][2] = r0.z, o[1][0] = r0.w, o[1][1] = r1.x, o[1][2] = r1.y, o[2][0] = r1.z, o[2][1] = r1.w, o[2][2] = r2.x;
		tc[0][0] = r2.y, tc[0][1] = r2.z, tc[1][0] = r2.w, tc[1][1] = r3.x, tc[2][0] = r3.y, tc[2][1] = r3.z;
		material_id = __float_as_uint(r3.w);
	} else {
#pragma unroll
		for (int k = 0; k < 3; ++k) { // GetSceneVertex / GetSceneTexcoord (Scene.glsl:50-56)
			const float *p = sc.vertices + 3 * (size_t)__ldg(sc.vertex_indices + 3 * (size_t)prim + k);
			o[k][0] = __ldg(p), o[k][1] = __ldg(p + 1), o[k][2] = __ldg(p + 2);
			const float2 t = __ldg((const float2 *)sc.texcoords + __ldg(sc.texcoord_indices + 3 * (size_t)prim + k));
			tc[k][0] = t.x, tc[k][1] = t.y;
		}
		material_id = __ldg(sc.material_ids + prim);
	}
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float x = o[k][0], y = o[k][1], z = o[k][2];
		v[k][0] = x * m0.x + y * m0.y + z * m0.z + m0.w;
		v[k][1] = x * m1.x + y * m1.y + z * m1.z + m1.w;
		v[k][2] = x * m2.x + y * m2.y + z * m2.z + m2.w;
	}
	const float e1x = v[1][0] - v[0][0], e1y = v[1][1] - v[0][1], e1z = v[1][2] - v[0][2];
	const float e2x = v[2][0] - v[0][0], e2y = v[2][1] - v[0][1], e2z = v[2][2] - v[0][2];
	float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
	const float inv = rsqrtf(nx * nx + ny * ny + nz * nz); // (2 ulp; the normal goes through atan2 / acos and a one-blob next)
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
	const NrcMaterial *mat = sc.materials + material_id;
	float4 md, ms; // diffuse + texture id, specular + texture id: the first 32 bytes of the 64-byte std430 element
	if (((uintptr_t)sc.materials & 31u) == 0u) // (uniform: a caller's buffer need only be 16-byte aligned)
		ldg256(mat, md, ms);
	else
		md = __ldg((const float4 *)mat), ms = __ldg((const float4 *)mat + 1);
	out[7] = __ldg(&mat->roughness);
	const float u = tc[0][0] * bx + tc[1][0] * by + tc[2][0] * bz, w = tc[0][1] * bx + tc[1][1] * by + tc[2][1] * bz;
	const uint32_t dtex = __float_as_uint(md.w), stex = __float_as_uint(ms.w);
	out[8] = md.x, out[9] = md.y, out[10] = md.z, out[11] = ms.x, out[12] = ms.y, out[13] = ms.z;
	if (dtex != 0xFFFFFFFFu) // GetSceneDiffuse / GetSceneSpecular (Scene.glsl:59-64)
		sample_texture_srgb_repeat(textures + dtex, u, w, out + 8, lut);
	if (stex != 0xFFFFFFFFu)
		sample_texture_srgb_repeat(textures + stex, u, w, out + 11, lut);
}

__device__ __forceinline__ void load_packed_input(const void *base, uint64_t index, uint32_t stride_bytes, uint32_t pk[4]) {
	const uint32_t *p = (const uint32_t *)((const uint8_t *)base + index * stride_bytes); // 4-byte aligned only (20 / 40 B records)
	pk[0] = __ldg(p), pk[1] = __ldg(p + 1), pk[2] = __ldg(p + 2), pk[3] = __ldg(p + 3);
}

} // namespace nrc
