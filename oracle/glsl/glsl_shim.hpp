// glsl_shim.hpp -- TEST INFRASTRUCTURE: the GLSL execution environment the reference's shaders need, in C++, so that the
// shader sources themselves (translated lexically by glsl2cpp.py, see the rule list there) compile with g++ and run on
// the CPU. Vector / matrix types and the GLSL built-in functions come from glm - the reference's own vendored copy
// (dep/glm/include) - wherever glm has them; this file adds what glm has not:
//   * float16_t and the cooperative-matrix extension (GL_NV_cooperative_matrix) as a functional emulation,
//   * compute built-in variables and barrier() (workgroups run as cooperative fibers, glsl_runtime.hpp),
//   * image2D / sampler2D with the sampler semantics the reference configures, buffer atomics, subgroupAll,
//   * the mixed int/float overloads GLSL's implicit conversions allow and C++ template deduction does not.
// What is NOT the reference's code and therefore remains a stated assumption of this emulation:
//   (E1) coopMatMulAddNV: D = fp16(C + sum_k A[i][k] * B[k][j]), products and the 16-term sum in fp32, one rounding to
//        fp16 per 16x16x16 MMA (the fp16-accumulator behaviour of the NV tensor-core path; the extension leaves the
//        internal precision to the implementation);
//   (E2) texture(): VK_FILTER_LINEAR per the Vulkan spec formula with fp32 weights, sRGB decoded per texel before the
//        blend (hardware uses 8-bit fixed-point weights: up to ~2e-3 of difference on a filtered colour);
//   (E3) packHalf2x16 / float16_t(x): round-to-nearest-even (what the GPU's cvt.rn.f16.f32 does; glm's own packHalf2x16
//        rounds ties up, so it is replaced);
//   (E4) transcendental built-ins (exp, atan, acos, inversesqrt, sqrt) are libm's, not the GPU's approximations.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define GLM_FORCE_SWIZZLE
#define _MSC_EXTENSIONS // glm/detail/setup.hpp:75: turns on the language-extension flag glm wants for `v.xy`-style swizzle members
#include <glm/glm.hpp>
#undef _MSC_EXTENSIONS
#undef M_PI // Constant.glsl defines its own

using namespace glm;
typedef _Float16 float16_t;
using f16vec3 = glm::vec<3, float16_t, glm::defaultp>;

// ---------------------------------------------------------------------------------------------------------- built-ins
extern thread_local uvec3 gl_LocalInvocationID, gl_GlobalInvocationID, gl_WorkGroupID;
extern thread_local uint gl_SubgroupID, gl_SubgroupInvocationID;
void barrier(); // glsl_runtime.cpp: switches to the next invocation of the workgroup
constexpr int gl_ScopeSubgroup = 3, gl_ScopeQueueFamily = 5, gl_StorageSemanticsBuffer = 0x40, gl_SemanticsRelaxed = 0;
inline bool subgroupAll(bool v) { return v; } // (only used on workgroup-uniform values, nrc_optimize.comp:33)

// ---------------------------------------------------------------------------------------------------------- fp16
inline float16_t uint16BitsToHalf(uint16_t b) {
	float16_t h;
	std::memcpy(&h, &b, 2);
	return h;
}
inline uint16_t glsl_half_bits(float16_t h) {
	uint16_t b;
	std::memcpy(&b, &h, 2);
	return b;
}
inline float16_t max(float16_t a, float16_t b) { return a > b ? a : b; }
inline bool isnan(float16_t a) { return a != a; }
inline uint glsl_packHalf2x16(const vec2 &v) { // (E3)
	return (uint)glsl_half_bits((float16_t)v.x) | ((uint)glsl_half_bits((float16_t)v.y) << 16);
}
inline vec2 glsl_unpackHalf2x16(uint u) { return vec2((float)uint16BitsToHalf((uint16_t)(u & 0xFFFFu)), (float)uint16BitsToHalf((uint16_t)(u >> 16))); }
#define packHalf2x16 glsl_packHalf2x16
#define unpackHalf2x16 glsl_unpackHalf2x16

// ---------------------------------------------------------------------------------------------------------- conversions GLSL has
inline float clamp(float x, int lo, int hi) { return glm::clamp(x, (float)lo, (float)hi); }
inline uint min(uint a, int b) { return a < (uint)b ? a : (uint)b; }
namespace glm {
template <length_t L, qualifier Q> inline vec<L, float, Q> operator-(int a, const vec<L, float, Q> &b) { return (float)a - b; }
template <length_t L, qualifier Q> inline vec<L, float, Q> operator+(int a, const vec<L, float, Q> &b) { return (float)a + b; }
template <length_t L, qualifier Q> inline vec<L, float, Q> operator*(int a, const vec<L, float, Q> &b) { return (float)a * b; }
namespace detail {
template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
inline bool operator==(const _swizzle<N, T, Q, E0, E1, E2, E3> &a, const vec<N, T, Q> &b) {
	return a() == b; // GLSL `==` on vectors: true iff all components are equal
}
} // namespace detail
} // namespace glm
template <class T, int N> struct glsl_array { // `T[N](...)` used as an expression (glsl2cpp.py R6)
	T v[N];
	operator const T *() const { return v; }
};

// ---------------------------------------------------------------------------------------------------------- cooperative matrices
// GL_NV_cooperative_matrix, emulated functionally: a matrix is subgroup-uniform, so every invocation of the subgroup simply
// holds the whole 16x16 matrix; length() is the whole matrix and the per-element loops of the shaders (ReLU, masks) apply
// the same function to every element, which is what they do on the GPU piecewise.
template <int Bits, int Scope, int Rows, int Cols> struct fcoopmatNV {
	static_assert(Bits == 16, "the reference only uses fp16 cooperative matrices");
	float16_t e[Rows * Cols]; // row-major
	fcoopmatNV() {}           // uninitialised, as in GLSL
	template <class S> explicit fcoopmatNV(S v) {
		for (int i = 0; i < Rows * Cols; ++i)
			e[i] = (float16_t)v;
	}
	uint length() const { return Rows * Cols; }
	float16_t &operator[](uint i) { return e[i]; }
	const float16_t &operator[](uint i) const { return e[i]; }
};
// element / stride are in units of the buffer's element type (uvec4 = 8 halfs); colMajor: column j is contiguous
template <int S, int R, int C> inline void coopMatLoadNV(fcoopmatNV<16, S, R, C> &m, const uvec4 *buf, uint element, uint stride, bool colMajor) {
	const float16_t *base = (const float16_t *)(buf + element);
	const size_t s = (size_t)stride * 8;
	for (int i = 0; i < R; ++i)
		for (int j = 0; j < C; ++j)
			m.e[i * C + j] = colMajor ? base[j * s + i] : base[i * s + j];
}
template <int S, int R, int C> inline void coopMatStoreNV(const fcoopmatNV<16, S, R, C> &m, uvec4 *buf, uint element, uint stride, bool colMajor) {
	float16_t *base = (float16_t *)(buf + element);
	const size_t s = (size_t)stride * 8;
	for (int i = 0; i < R; ++i)
		for (int j = 0; j < C; ++j)
			(colMajor ? base[j * s + i] : base[i * s + j]) = m.e[i * C + j];
}
// (E1); the arithmetic lives in glsl_runtime.cpp together with the per-subgroup memo that lets 31 of the 32 invocations
// reuse the product the first one computed (identical operands by construction)
void glsl_coopmat_muladd_16(const float16_t *a, const float16_t *b, const float16_t *c, float16_t *d);
template <int S> inline fcoopmatNV<16, S, 16, 16> coopMatMulAddNV(const fcoopmatNV<16, S, 16, 16> &a, const fcoopmatNV<16, S, 16, 16> &b,
                                                                   const fcoopmatNV<16, S, 16, 16> &c) {
	fcoopmatNV<16, S, 16, 16> d;
	glsl_coopmat_muladd_16(a.e, b.e, c.e, d.e);
	return d;
}

// ---------------------------------------------------------------------------------------------------------- buffers, images, samplers
void glsl_atomic_add(float *p, float v);
inline void atomicAdd(float &mem, float v, int, int, int) { glsl_atomic_add(&mem, v); }

enum glsl_format { GLSL_RGBA32F = 0, GLSL_RG32F = 1, GLSL_RGBA8 = 2 };
struct image2D {
	void *data = nullptr;
	int width = 0, height = 0, pitch = 0; // pitch in pixels
	int format = GLSL_RGBA32F;
};
inline vec4 imageLoad(const image2D &im, ivec2 c) {
	const size_t at = (size_t)c.y * im.pitch + c.x;
	if (im.format == GLSL_RGBA32F) {
		const float *p = (const float *)im.data + 4 * at;
		return vec4(p[0], p[1], p[2], p[3]);
	}
	const float *p = (const float *)im.data + 2 * at; // rg32f: missing components read (0, 1)
	return vec4(p[0], p[1], 0.0f, 1.0f);
}
inline void imageStore(image2D &im, ivec2 c, vec4 v) {
	const size_t at = (size_t)c.y * im.pitch + c.x;
	if (im.format == GLSL_RGBA32F) {
		float *p = (float *)im.data + 4 * at;
		p[0] = v.x, p[1] = v.y, p[2] = v.z, p[3] = v.w;
	} else if (im.format == GLSL_RGBA8) { // float -> unorm8: clamp, scale, round to nearest even
		uint8_t *p = (uint8_t *)im.data + 4 * at;
		for (int k = 0; k < 4; ++k)
			p[k] = (uint8_t)std::nearbyintf(glm::clamp(v[k], 0.0f, 1.0f) * 255.0f);
	}
}
struct sampler2D { // VK_FILTER_LINEAR, one mip level; sRGB or UNORM RGBA8 texels
	const uint8_t *texels = nullptr;
	int width = 0, height = 0;
	bool srgb = true;   // scene textures: VK_FORMAT_R8G8B8A8_SRGB (src/VkScene.cpp:193)
	bool repeat = true; // scene: ADDRESS_MODE_REPEAT (src/rg/NRCRenderGraph.cpp:132); learn-an-image: CLAMP_TO_EDGE (main.cpp:121-124)
};
vec4 texture(const sampler2D &s, vec2 uv); // (E2), glsl_runtime.cpp
