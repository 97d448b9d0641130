// nrc_encode.cuh -- input encodings, evaluated per thread (= per sample) and packed straight into the fp16x2
// registers that become the first layer's A operand.
//   NRCInputEncode        shader/src/NRCRecord.glsl:47-95
//   one-blob-32 (image)   test/mlp_learning_an_image/gradient.comp:26-44
// Arithmetic contract (tests/test_gpu_parity.py::test_fused_encoding_*):
//   * frequency features: BIT-EXACT against the unfused fp32 evaluation of the GLSL source (oracle/nrc_oracle.c); the
//     rewrite below only uses steps that are exact in fp32 (see tri()).
//   * one-blob features: the quartic kernel is evaluated in Horner form with FMAs (GLSL leaves contraction to the
//     compiler - the source carries no `precise` - so the reference's own last bit is not defined); the result is within
//     2 fp32 ulps of the unfused evaluation, i.e. at most one fp16 ulp (or 2.4e-7 near zero) on the encoded feature.
#pragma once
#include "sm100_ptx.cuh"

namespace nrc {

// _quartic_cdf(e - x, inv_radius) (NRCRecord.glsl:51-56) for a power-of-two inv_radius: u = fl(fl(e - x) * inv_radius)
// == fl(e * inv_radius - x * inv_radius) (scaling by a power of two commutes with rounding) = one FMA; 6 instructions.
__device__ __forceinline__ float quartic_cdf_edge(float x, float edge, float inv_radius) {
	const float u = fmaf(x, -inv_radius, edge * inv_radius);
	const float u2 = u * u;
	const float poly = fmaf(u2, fmaf(u2, 1.0f / 5.0f, -2.0f / 3.0f), 1.0f);
	return __saturatef(fmaf(u * poly, 15.0f / 16.0f, 0.5f));
}

// NRCOneBlob4Encode (NRCRecord.glsl:58-63). Bin i is cdf(r_i - x) - cdf(l_i - x) with r_i == l_{i+1}, so the five
// distinct edge values are evaluated once each.
__device__ __forceinline__ void oneblob4(float x, float out[4]) {
	float c[5];
#pragma unroll
	for (int i = 0; i < 5; ++i)
		c[i] = quartic_cdf_edge(x, 0.25f * (float)i, 4.0f);
#pragma unroll
	for (int i = 0; i < 4; ++i)
		out[i] = c[i + 1] - c[i];
}

// _nrc_tri (NRCRecord.glsl:65-68) of x = scale * p (scale a power of two): 2|mod(x - 1/2, 2) - 1| - 1, in 5 instructions,
// bit-identical to the step-by-step evaluation:  a = fl(x - 1/2) is the only rounding of the source. h = a / 2 =
// fl(scale/2 * p - 1/4) (same rounding, scaled); mod(a, 2) = 2 (h - floor h) with h - floor(h) exact; the source's
// fl(m - 1) equals fl(4 frac - 2) / 2 (again a scaled rounding) and 2|.| - 1 is exact.
__device__ __forceinline__ float tri_scaled(float p, float scale) {
	const float h = fmaf(p, 0.5f * scale, -0.25f);
	const float frac = h - floorf(h);
	return fabsf(fmaf(frac, 4.0f, -2.0f)) - 1.0f;
}

// ---- packed-fp32 forms. Blackwell's fma.rn.f32x2 / mul.rn.f32x2 perform two IEEE fp32 operations per instruction,
// each bit-identical to its scalar counterpart, so everything stated above (bit-exact frequency features, one-blob
// features within one fp16 ulp) holds unchanged; the encoder drops from ~400 to ~300 arithmetic instructions per record.
#ifndef NRC_ENCODE_SCALAR
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// tri_scaled(p, s0), tri_scaled(p, s1)
__device__ __forceinline__ float2 tri_scaled_x2(float p, float s0, float s1) {
	const float2 h = __ffma2_rn(splat2(p), make_float2(0.5f * s0, 0.5f * s1), splat2(-0.25f));
	const float2 fl = make_float2(floorf(h.x), floorf(h.y));
	const float2 frac = __ffma2_rn(fl, splat2(-1.0f), h); // h - floor(h): the product is exact, one rounding as in the scalar form
	const float2 y = __ffma2_rn(frac, splat2(4.0f), splat2(-2.0f));
	return make_float2(fabsf(y.x) - 1.0f, fabsf(y.y) - 1.0f);
}
// quartic_cdf_edge(x.x, e.x / r, r), quartic_cdf_edge(x.y, e.y / r, r) with e = edge * inv_radius (exact: power-of-two radii)
__device__ __forceinline__ float2 quartic_cdf_edge_x2(float2 x, float2 edge_scaled, float inv_radius) {
	const float2 u = __ffma2_rn(x, splat2(-inv_radius), edge_scaled);
	const float2 u2 = __fmul2_rn(u, u);
	const float2 poly = __ffma2_rn(u2, __ffma2_rn(u2, splat2(1.0f / 5.0f), splat2(-2.0f / 3.0f)), splat2(1.0f));
	const float2 r = __ffma2_rn(__fmul2_rn(u, poly), splat2(15.0f / 16.0f), splat2(0.5f));
	return make_float2(__saturatef(r.x), __saturatef(r.y));
}
// oneblob4 of two inputs at once
__device__ __forceinline__ void oneblob4_x2(float a, float b, float outa[4], float outb[4]) {
	float2 c[5];
#pragma unroll
	for (int i = 0; i < 5; ++i)
		c[i] = quartic_cdf_edge_x2(make_float2(a, b), splat2(0.25f * (float)i * 4.0f), 4.0f);
#pragma unroll
	for (int i = 0; i < 4; ++i)
		outa[i] = c[i + 1].x - c[i].x, outb[i] = c[i + 1].y - c[i].y;
}
// tri_scaled(p, 2^k) for k = k0 .. k0 + count - 1 (count even)
template <int k0, int count> __device__ __forceinline__ void tri_octaves(float p, float *f) {
#pragma unroll
	for (int k = 0; k < count; k += 2) {
		const float2 t = tri_scaled_x2(p, (float)(1 << (k0 + k)), (float)(2 << (k0 + k)));
		f[k] = t.x, f[k + 1] = t.y;
	}
}
#else
__device__ __forceinline__ void oneblob4_x2(float a, float b, float outa[4], float outb[4]) { oneblob4(a, outa), oneblob4(b, outb); }
template <int k0, int count> __device__ __forceinline__ void tri_octaves(float p, float *f) {
#pragma unroll
	for (int k = 0; k < count; ++k)
		f[k] = tri_scaled(p, (float)(1 << (k0 + k)));
}
#endif

// 14 floats (UnpackedNRCInput order) -> 64 features as 32 packed fp16 pairs, slot order NRCRecord.glsl:86-94.
__device__ __forceinline__ void encode_nrc(const float in[14], uint32_t o[32]) {
	float f[64];
#pragma unroll
	for (int a = 0; a < 3; ++a)
		tri_octaves<0, 12>(in[a], f + 12 * a);
	oneblob4_x2(in[3], in[4], f + 36, f + 40);
	oneblob4_x2(in[5], in[6], f + 44, f + 48);
	oneblob4(1.0f - expf(-in[7]), f + 52);
#pragma unroll
	for (int i = 0; i < 6; ++i)
		f[56 + i] = in[8 + i];
	f[62] = 1.0f, f[63] = 1.0f;
#pragma unroll
	for (int i = 0; i < 32; ++i)
		o[i] = sm100::cvt_pack_f16x2(f[2 * i], f[2 * i + 1]);
}

// The same encoding split in two 32-feature halves (half 0 = features 0..31, half 1 = features 32..63), so that two
// threads can share one record (the training kernel's epilogue warps each own 32 of a row's 64 columns).
__device__ __forceinline__ void encode_nrc_half(const float in[14], uint32_t half, uint32_t o[16]) {
	float f[32];
	if (half == 0) {
		tri_octaves<0, 12>(in[0], f);
		tri_octaves<0, 12>(in[1], f + 12);
		tri_octaves<0, 8>(in[2], f + 24);
	} else {
		tri_octaves<8, 4>(in[2], f);
		oneblob4_x2(in[3], in[4], f + 4, f + 8);
		oneblob4_x2(in[5], in[6], f + 12, f + 16);
		oneblob4(1.0f - expf(-in[7]), f + 20);
#pragma unroll
		for (int i = 0; i < 6; ++i)
			f[24 + i] = in[8 + i];
		f[30] = 1.0f, f[31] = 1.0f;
	}
#pragma unroll
	for (int i = 0; i < 16; ++i)
		o[i] = sm100::cvt_pack_f16x2(f[2 * i], f[2 * i + 1]);
}

// one-blob-32 of a single coordinate (32 features): gradient.comp:33-39 incl. the mismatched inverse radii 32 / 4
__device__ __forceinline__ void encode_oneblob32_half(float x, uint32_t o[16]) {
#pragma unroll
	for (int i = 0; i < 16; ++i) {
		const float l0 = (float)(2 * i) / 32.0f, r0 = (float)(2 * i + 1) / 32.0f, r1 = (float)(2 * i + 2) / 32.0f;
		const float e0 = quartic_cdf_edge(x, r0, 32.0f) - quartic_cdf_edge(x, l0, 4.0f);
		const float e1 = quartic_cdf_edge(x, r1, 32.0f) - quartic_cdf_edge(x, r0, 4.0f);
		o[i] = sm100::cvt_pack_f16x2(e0, e1);
	}
}

// one-blob-32 of u then v; note the mismatched inverse radii 32 / 4 are the reference's (gradient.comp:33-39).
__device__ __forceinline__ void encode_oneblob32(float u, float v, uint32_t o[32]) {
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const float x = h ? v : u;
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			const float l0 = (float)(2 * i) / 32.0f, r0 = (float)(2 * i + 1) / 32.0f, r1 = (float)(2 * i + 2) / 32.0f;
			const float e0 = quartic_cdf_edge(x, r0, 32.0f) - quartic_cdf_edge(x, l0, 4.0f);
			const float e1 = quartic_cdf_edge(x, r1, 32.0f) - quartic_cdf_edge(x, r0, 4.0f);
			o[16 * h + i] = sm100::cvt_pack_f16x2(e0, e1);
		}
	}
}

// pcg2d (test/mlp_learning_an_image/gradient.comp:15-24)
__device__ __forceinline__ void pcg2d(uint32_t &x, uint32_t &y) {
	x = x * 1664525u + 1013904223u;
	y = y * 1664525u + 1013904223u;
	x += y * 1664525u;
	y += x * 1664525u;
	x ^= x >> 16;
	y ^= y >> 16;
	x += y * 1664525u;
	y += x * 1664525u;
	x ^= x >> 16;
	y ^= y >> 16;
}

} // namespace nrc
