/*
 * nrc_oracle.c -- CPU restatement of VkNRC's Neural-Radiance-Cache MLP hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in vknrc_b200/ (the product) may import, link or execute this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: the reference ships NO golden vectors, fixtures or thresholds; this oracle is pinned by
 * reference code executed here: (1) its forward pass against the reference's own CPU `Evaluate` (test/main.cpp:11-27)
 * compiled unmodified into oracle/_ref/ (tests/test_oracle.py, tests/golden/nrc_golden_v1.npz); (2) every other function -
 * encoding, scene gather, dst codec, loss gradients, backward pass, dW reduction, prepare + optimizer, scatter, the
 * learn-an-image kernels - against the reference's own GLSL SHADERS compiled as C++ and run on the CPU (oracle/glsl ->
 * oracle/_ref_glsl, tests/golden/nrc_golden_v2.npz, tests/test_ref_glsl.py): bit-exact for encoding, forward in shader
 * precision, optimizer, dst codec; to the order of the shader's fp32 atomics for the gradients. What stays an assumption
 * is how the GPU's cooperative-matrix MMA and texture unit round internally (E1, E2 in oracle/glsl/glsl_shim.hpp): the
 * GLSL cannot be executed on a GPU here (no Vulkan loader / ICD).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * Build: see oracle/Makefile (gcc -O2 -mf16c -ffp-contract=off: no FMA contraction, so fp32 steps round
 * exactly where the GLSL source rounds).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 f16;

#define NRC_WIDTH 64
#define NRC_OUT 3
#define NRC_HIDDEN 5
#define NRC_WEIGHTS (NRC_WIDTH * NRC_WIDTH * NRC_HIDDEN + NRC_WIDTH * NRC_OUT) /* src/VkNRCState.hpp:23-25 */

/* accumulate modes */
#define ACC_FP16_CHUNK16 0 /* shader-like: fp16 accumulator matrix, K consumed in 16-wide MMAs (NN_nv.glsl:106-117) */
#define ACC_FP32 1         /* B200 kernel-like: fp32 accumulate over all of K, round once to fp16 */

/* loss kinds */
#define LOSS_L2 0          /* NN_nv.glsl:162-177 (test/train_NV.comp, learn-an-image) */
#define LOSS_RELATIVE_L2_LUMINANCE 1 /* NN_nv.glsl:178-196 (nrc_gradient.comp) */

static inline f16 u2h(uint16_t u) { f16 h; memcpy(&h, &u, 2); return h; }
static inline uint16_t h2u(f16 h) { uint16_t u; memcpy(&u, &h, 2); return u; }

/* ------------------------------------------------------------------------------------------------------------
 * Encoding  (shader/src/NRCRecord.glsl:47-95)
 * ---------------------------------------------------------------------------------------------------------- */

/* NRCRecord.glsl:51-56 `_quartic_cdf`, and test/mlp_learning_an_image/gradient.comp:26-31 */
static float quartic_cdf(float x, float inv_radius) {
	float u = x * inv_radius;
	float u2 = u * u;
	float u4 = u2 * u2;
	float poly = (1.0f - (2.0f / 3.0f) * u2) + (1.0f / 5.0f) * u4;
	float v = ((15.0f / 16.0f) * u) * poly + 0.5f;
	return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
}

/* NRCRecord.glsl:58-63 `NRCOneBlob4Encode` */
static void oneblob4(float x, float out[4]) {
	static const float l[4] = {0.0f, 0.25f, 0.5f, 0.75f}, r[4] = {0.25f, 0.5f, 0.75f, 1.0f};
	for (int i = 0; i < 4; ++i)
		out[i] = quartic_cdf(r[i] - x, 4.0f) - quartic_cdf(l[i] - x, 4.0f);
}

/* NRCRecord.glsl:65-68 `_nrc_tri`; GLSL mod(x, y) = x - y * floor(x / y) */
static float nrc_tri(float x) {
	float a = x - 0.5f;
	float m = a - 2.0f * floorf(a / 2.0f);
	return 2.0f * fabsf(m - 1.0f) - 1.0f;
}

/* NRCRecord.glsl:69-72 `NRCFrequencyEncode`: 12 octaves 2^k * p, k = 0..11 */
static void freq12(float p, float out[12]) {
	for (int k = 0; k < 12; ++k)
		out[k] = nrc_tri((float)(1 << k) * p);
}

/* NRCRecord.glsl:78-95 `NRCInputEncode`. `in14` is UnpackedNRCInput flattened in declaration order
 * (NRCRecord.glsl:40-45): position.xyz, scattered_dir.xy, normal.xy, roughness, diffuse.rgb, specular.rgb.
 * Output slot order follows o[0..7] at :86-94. fp32 -> fp16 is round-to-nearest-even (packHalf2x16). */
void nrc_oracle_encode(const float *in14, uint16_t *out64) {
	float f[64];
	freq12(in14[0], f + 0);
	freq12(in14[1], f + 12);
	freq12(in14[2], f + 24);
	oneblob4(in14[3], f + 36);
	oneblob4(in14[4], f + 40);
	oneblob4(in14[5], f + 44);
	oneblob4(in14[6], f + 48);
	oneblob4(1.0f - expf(-in14[7]), f + 52);
	f[56] = in14[8], f[57] = in14[9], f[58] = in14[10];
	f[59] = in14[11], f[60] = in14[12], f[61] = in14[13];
	f[62] = 1.0f, f[63] = 1.0f;
	for (int i = 0; i < 64; ++i)
		out64[i] = h2u((f16)f[i]);
}
void nrc_oracle_encode_batch(const float *in14, uint64_t n, uint16_t *out64) {
	for (uint64_t i = 0; i < n; ++i)
		nrc_oracle_encode(in14 + 14 * i, out64 + 64 * i);
}

/* test/mlp_learning_an_image/gradient.comp:33-44 (`oneblob_32`, `pack_half_32`) and inference.comp:19-30.
 * NOTE the mismatched inverse radii 32 / 4 are the reference's and are reproduced as written. */
void nrc_oracle_encode_oneblob32(float u, float v, uint16_t *out64) {
	for (int i = 0; i < 32; ++i) {
		float l = (float)i / 32.0f, r = (float)(i + 1) / 32.0f;
		out64[i] = h2u((f16)(quartic_cdf(r - u, 32.0f) - quartic_cdf(l - u, 4.0f)));
		out64[32 + i] = h2u((f16)(quartic_cdf(r - v, 32.0f) - quartic_cdf(l - v, 4.0f)));
	}
}

/* test/mlp_learning_an_image/gradient.comp:15-24 `pcg2d` and :47-48 (uv of sample `gid`) */
void nrc_oracle_pcg2d(uint32_t x, uint32_t y, uint32_t *ox, uint32_t *oy) {
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
	*ox = x, *oy = y;
}
void nrc_oracle_learn_image_uv(uint32_t seed_x, uint32_t seed_y, uint32_t gid, float *u, float *v) {
	uint32_t px, py;
	nrc_oracle_pcg2d(seed_x + gid % 128u, seed_y + gid / 128u, &px, &py);
	const float s = 1.0f / (float)0xffffffffu;
	*u = s * (float)px;
	*v = s * (float)py;
}

/* ------------------------------------------------------------------------------------------------------------
 * Eval-record destination codec  (shader/src/NRCRecord.glsl:12-33) -- integer work, must be bit-exact
 * ---------------------------------------------------------------------------------------------------------- */
uint32_t nrc_oracle_dst_screen(uint32_t x15, uint32_t y15) { return (x15 | (y15 << 15)) << 1; }
uint32_t nrc_oracle_dst_train(uint32_t b2, uint32_t l14, uint32_t r14) {
	return ((b2 | (l14 << 2) | (r14 << 16)) << 1) | 1u;
}
void nrc_oracle_dst_decode(uint32_t e, uint32_t *type, uint32_t *a, uint32_t *b, uint32_t *c) {
	*type = e & 1u;
	e >>= 1;
	if (*type == 0u) {
		*a = e & 0x7FFFu, *b = e >> 15, *c = 0;
	} else {
		*a = e & 3u, *b = (e >> 2) & 0x3FFFu, *c = e >> 16;
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * Forward  (shader/src/NN_nv.glsl:84-158; CPU twin test/main.cpp:11-27)
 * Weights: fp16 row-major W[l][out][in], layer l at l*4096, layer 5 = 3x64 at 20480 (NN_nv.glsl:69-82).
 * Activations are sample-major [n][64]. `acts` (optional) receives a_0..a_5, each n*64 fp16, back to back.
 * `y` receives the 3 linear outputs per sample, widened fp16 -> fp32 exactly (NN_nv.glsl:148-158).
 * ---------------------------------------------------------------------------------------------------------- */
static f16 dot64(const uint16_t *wrow, const f16 *a, int mode) {
	if (mode == ACC_FP32) {
		float acc = 0.0f;
		for (int i = 0; i < 64; ++i)
			acc += (float)u2h(wrow[i]) * (float)a[i];
		return (f16)acc;
	}
	f16 acc = (f16)0.0f; /* fp16 accumulator, one 16-wide MMA at a time (NN_nv.glsl:110-117) */
	for (int c = 0; c < 4; ++c) {
		float s = 0.0f; /* intra-MMA precision is hardware-defined; fp32 here */
		for (int i = 16 * c; i < 16 * c + 16; ++i)
			s += (float)u2h(wrow[i]) * (float)a[i];
		acc = (f16)((float)acc + s);
	}
	return acc;
}

static void forward_one(const uint16_t *w, const uint16_t *x, int mode, f16 act[6][64], f16 y[3]) {
	for (int i = 0; i < 64; ++i)
		act[0][i] = u2h(x[i]);
	for (int l = 0; l < NRC_HIDDEN; ++l) { /* NNForward64_ReLU, NN_nv.glsl:99-127 */
		const uint16_t *wl = w + l * 4096;
		for (int o = 0; o < 64; ++o) {
			f16 z = dot64(wl + o * 64, act[l], mode);
			act[l + 1][o] = z > (f16)0.0f ? z : (f16)0.0f;
		}
	}
	for (int o = 0; o < NRC_OUT; ++o) /* NNForward3 + NNOutput3, NN_nv.glsl:129-158: no activation */
		y[o] = dot64(w + 5 * 4096 + o * 64, act[5], mode);
}

void nrc_oracle_forward(const uint16_t *w, const uint16_t *x, uint64_t n, int mode, uint16_t *acts, float *y) {
#pragma omp parallel for schedule(static)
	for (int64_t s = 0; s < (int64_t)n; ++s) {
		f16 act[6][64], yo[3];
		forward_one(w, x + 64 * s, mode, act, yo);
		if (acts)
			for (int l = 0; l < 6; ++l)
				for (int i = 0; i < 64; ++i)
					acts[(uint64_t)l * n * 64 + 64 * s + i] = h2u(act[l][i]);
		for (int o = 0; o < 3; ++o)
			y[3 * s + o] = (float)yo[o];
	}
}

/* Inference output as the test kernel stores it: F16Vec3 per sample (test/evaluate_NV.comp:29-30). `clamp` != 0
 * applies nrc_inference.comp:48 `max(predict, 0)`. */
void nrc_oracle_evaluate(const uint16_t *w, const uint16_t *x, uint64_t n, int mode, int clamp, uint16_t *out) {
	float *y = (float *)malloc(sizeof(float) * 3 * n);
	nrc_oracle_forward(w, x, n, mode, NULL, y);
	for (uint64_t i = 0; i < 3 * n; ++i) {
		float v = y[i];
		if (clamp && !(v > 0.0f))
			v = 0.0f;
		out[i] = h2u((f16)v);
	}
	free(y);
}

/* ------------------------------------------------------------------------------------------------------------
 * Loss gradient  (NN_nv.glsl:162-196), fp32 then rounded to fp16 by packHalf2x16
 * ---------------------------------------------------------------------------------------------------------- */
static void loss_grad(const float y[3], const float t[3], int kind, float loss_scale, f16 g[3]) {
	if (kind == LOSS_L2) { /* :165 `2.0 * (predict - target) * loss_scale` */
		for (int c = 0; c < 3; ++c)
			g[c] = (f16)(2.0f * (y[c] - t[c]) * loss_scale);
	} else { /* :183-184 */
		float p0 = y[0] > 0.0f ? y[0] : 0.0f, p1 = y[1] > 0.0f ? y[1] : 0.0f, p2 = y[2] > 0.0f ? y[2] : 0.0f;
		float lum = 0.299f * p0 + 0.587f * p1 + 0.114f * p2;
		float den = lum * lum + 0.01f;
		for (int c = 0; c < 3; ++c)
			g[c] = (f16)(2.0f * loss_scale * (y[c] - t[c]) / den);
	}
}

/* Scalar loss (never computed by the reference; SURVEY A.4 defines it for loss-curve parity):
 * L2: mean_n sum_c (y-t)^2;  relative: mean_n sum_c (y-t)^2 / (lum(max(y,0))^2 + 0.01). Double accumulation. */
double nrc_oracle_loss(const float *y, const float *t, uint64_t n, int kind) {
	double acc = 0.0;
	for (uint64_t s = 0; s < n; ++s) {
		const float *ys = y + 3 * s, *ts = t + 3 * s;
		double den = 1.0;
		if (kind == LOSS_RELATIVE_L2_LUMINANCE) {
			double lum = 0.299 * fmax(ys[0], 0.0) + 0.587 * fmax(ys[1], 0.0) + 0.114 * fmax(ys[2], 0.0);
			den = lum * lum + 0.01;
		}
		for (int c = 0; c < 3; ++c) {
			double d = (double)ys[c] - (double)ts[c];
			acc += d * d / den;
		}
	}
	return n ? acc / (double)n : 0.0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Backward + weight gradient  (nrc_gradient.comp:36-57 / test/train_NV.comp:21-45, NN_nv.glsl:198-367)
 *
 * delta_5 = g (3 values); for l = 5..0: dW_l += delta_l (x) a_l ; delta_{l-1} = (W_l^T delta_l) * [a_l > 0], l >= 1.
 * mode ACC_FP16_CHUNK16 follows the shader's rounding: fp16 accumulators for dA (NaN-sentinel mask == multiply by
 * [a>0], NN_nv.glsl:198-220,271-277), per-subgroup (32 samples) dW partial held in fp16 and accumulated one
 * 16-sample MMA at a time (NN_nv.glsl:284-289, 323-332), then fp32 across subgroups/workgroups (:297-316).
 * mode ACC_FP32 is what the sm_100a kernel does: fp32 accumulate, deltas rounded once to fp16, dW summed in fp32
 * (here in double so that the oracle is the better-rounded side).
 * `dw` ACCUMULATES (like the reference's atomicAdd into uDWeights): caller clears it.
 * Samples beyond n in the last 128-group contribute exactly zero in the reference (nrc_gradient.comp:27-34), so
 * looping over n samples only is equivalent.
 * ---------------------------------------------------------------------------------------------------------- */
void nrc_oracle_gradient(const uint16_t *w, const uint16_t *x, const float *targets, uint64_t n, int loss_kind,
                         float loss_scale, int mode, float *dw, float *y_out) {
	double *acc = (double *)calloc(NRC_WEIGHTS, sizeof(double));
	const uint64_t groups = (n + 31) / 32; /* one subgroup = 32 samples (NN_nv.glsl:52) */
	f16(*part)[64] = (f16(*)[64])malloc(sizeof(f16) * 64 * 64);
	for (uint64_t g = 0; g < groups; ++g) {
		const uint64_t s0 = g * 32, s1 = (s0 + 32 < n) ? s0 + 32 : n;
		static f16 act[32][6][64];
		static f16 delta[32][6][64]; /* delta[s][l] = dL/dz of layer l's output; layer 5 uses first 3 */
		for (uint64_t s = s0; s < s1; ++s) {
			f16 yo[3];
			float yf[3];
			forward_one(w, x + 64 * s, mode, act[s - s0], yo);
			for (int c = 0; c < 3; ++c)
				yf[c] = (float)yo[c];
			if (y_out)
				memcpy(y_out + 3 * s, yf, sizeof(yf));
			f16 gl[3];
			loss_grad(yf, targets + 3 * s, loss_kind, loss_scale, gl);
			f16(*d)[64] = delta[s - s0];
			memset(d, 0, sizeof(f16) * 6 * 64);
			for (int c = 0; c < 3; ++c)
				d[5][c] = gl[c];
			/* NNBackwardDA3_ReLU (l = 5, K = 16 padded; one MMA) then NNBackwardDA64_ReLU (l = 4..1) */
			for (int l = 5; l >= 1; --l) {
				const int outs = (l == 5) ? 3 : 64;
				const uint16_t *wl = w + l * 4096;
				for (int i = 0; i < 64; ++i) {
					float r;
					if (mode == ACC_FP32) {
						float a32 = 0.0f;
						for (int o = 0; o < outs; ++o)
							a32 += (float)d[l][o] * (float)u2h(wl[o * 64 + i]);
						r = a32;
					} else {
						f16 a16 = (f16)0.0f;
						for (int c0 = 0; c0 < outs; c0 += 16) {
							float sacc = 0.0f;
							for (int o = c0; o < c0 + 16 && o < outs; ++o)
								sacc += (float)d[l][o] * (float)u2h(wl[o * 64 + i]);
							a16 = (f16)((float)a16 + sacc);
						}
						r = (float)a16;
					}
					f16 v = (f16)r;
					if (!(act[s - s0][l][i] > (f16)0.0f) || v != v) /* mask on post-ReLU a_l; NaN -> 0 (Q3) */
						v = (f16)0.0f;
					d[l - 1][i] = v;
				}
			}
		}
		/* NNUpdateDW3 / NNUpdateDW64 for this subgroup */
		for (int l = 5; l >= 0; --l) {
			const int outs = (l == 5) ? 3 : 64;
			if (mode == ACC_FP32) {
				for (uint64_t s = s0; s < s1; ++s)
					for (int o = 0; o < outs; ++o) {
						const double dv = (double)(float)delta[s - s0][l][o];
						if (dv == 0.0)
							continue;
						for (int i = 0; i < 64; ++i)
							acc[l * 4096 + o * 64 + i] += dv * (double)(float)act[s - s0][l][i];
					}
			} else {
				for (int o = 0; o < outs; ++o)
					for (int i = 0; i < 64; ++i)
						part[o][i] = (f16)0.0f;
				for (uint64_t c0 = s0; c0 < s1; c0 += 16) /* one 16-sample MMA at a time */
					for (int o = 0; o < outs; ++o)
						for (int i = 0; i < 64; ++i) {
							float sacc = 0.0f;
							for (uint64_t s = c0; s < c0 + 16 && s < s1; ++s)
								sacc += (float)delta[s - s0][l][o] * (float)act[s - s0][l][i];
							part[o][i] = (f16)((float)part[o][i] + sacc);
						}
				for (int o = 0; o < outs; ++o)
					for (int i = 0; i < 64; ++i)
						acc[l * 4096 + o * 64 + i] += (double)(float)part[o][i];
			}
		}
	}
	for (int i = 0; i < NRC_WEIGHTS; ++i)
		dw[i] += (float)acc[i];
	free(part);
	free(acc);
}

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer  (shader/src/nrc_train_prepare.comp:16-28, shader/src/nrc_optimize.comp:32-54)
 * Layouts mirror src/VkNRCState.cpp:25-31.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { uint32_t t; float beta1_t, beta2_t, alpha_t, alpha_t_1; } NrcOracleOptimizerState;
typedef struct { float m, v, weight, ema_weight; } NrcOracleOptimizerEntry;

#define NRC_TRAIN_BATCH_SIZE 16384u /* shader/src/Constant.glsl:7 */
#define ADAM_BETA1 0.9f
#define ADAM_BETA2 0.999f
#define EMA_ALPHA 0.99f
#define LEARNING_RATE 0.002f /* nrc_optimize.comp:10 */
#define ADAM_EPSILON 1e-8f   /* nrc_optimize.comp:11 */
#define LOSS_SCALE 1.0f      /* Constant.glsl:8 */

/* nrc_train_prepare.comp:16-28. Returns the clamped count. `batch_cap` = 16384 in the reference. */
uint32_t nrc_oracle_prepare(uint32_t count, uint32_t batch_cap, NrcOracleOptimizerState *st) {
	if (count > batch_cap)
		count = batch_cap;
	if (count > 0) {
		++st->t;
		st->beta1_t *= ADAM_BETA1;
		st->beta2_t *= ADAM_BETA2;
		st->alpha_t_1 = st->alpha_t;
		st->alpha_t *= EMA_ALPHA;
	}
	return count;
}

/* nrc_optimize.comp:32-54. `use_weights` may be NULL (the non-WRITE_USE_WEIGHTS variant). */
void nrc_oracle_optimize(uint32_t count, const NrcOracleOptimizerState *st, NrcOracleOptimizerEntry *entries,
                         const float *gradients, uint16_t *weights, uint16_t *use_weights, int use_ema) {
	if (count == 0)
		return; /* :33-34 */
	for (int i = 0; i < NRC_WEIGHTS; ++i) {
		float g = gradients[i] / (float)count / LOSS_SCALE;
		if (isnan(g) || isinf(g))
			g = 0.0f;
		NrcOracleOptimizerEntry e = entries[i];
		e.m = ADAM_BETA1 * e.m + (1.0f - ADAM_BETA1) * g;
		e.v = ADAM_BETA2 * e.v + (1.0f - ADAM_BETA2) * (g * g);
		float hm = e.m / (1.0f - st->beta1_t), hv = e.v / (1.0f - st->beta2_t);
		e.weight -= LEARNING_RATE * hm / (sqrtf(hv) + ADAM_EPSILON);
		float eta_t = 1.0f - st->alpha_t, eta_t_1 = 1.0f - st->alpha_t_1;
		e.ema_weight = (1.0f - EMA_ALPHA) / eta_t * e.weight + EMA_ALPHA * eta_t_1 * e.ema_weight; /* sic, Q6 */
		entries[i] = e;
		weights[i] = h2u((f16)e.weight);
		if (use_weights)
			use_weights[i] = h2u((f16)(use_ema ? e.ema_weight : e.weight));
	}
}

/* test/mlp_learning_an_image/optimize.comp:21-29: SGD lr 0.01, gradient / 16384, skip on NaN/Inf. */
void nrc_oracle_sgd(float *fp_weights, const float *gradients, uint16_t *weights, float lr, float batch) {
	for (int i = 0; i < NRC_WEIGHTS; ++i) {
		float g = gradients[i] / batch / 1.0f;
		if (isnan(g) || isinf(g))
			continue;
		float wv = fp_weights[i];
		wv -= lr * g;
		fp_weights[i] = wv;
		weights[i] = h2u((f16)wv);
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * Inference scatter  (shader/src/nrc_inference.comp:48-73)
 * predict: n*3 fp32 (already max(.,0)); dst: n u32. Screen: rgba32f image `bias_factor_r` (pitch in pixels) RMW
 * plus rg32f `factor_gb`. Train: records[b][l..r].bias += factor * predict. Train records are 10 floats-worth:
 * {bias rgb, factor rgb, 4 x u32 packed input} = 40 B (NRCRecord.glsl:35-38).
 * ---------------------------------------------------------------------------------------------------------- */
void nrc_oracle_scatter(const float *predict, const uint32_t *dst, uint64_t n, float *bias_factor_r,
                        const float *factor_gb, uint32_t width, float *train_records[4]) {
	for (uint64_t s = 0; s < n; ++s) {
		uint32_t e = dst[s];
		if (e == 0xFFFFFFFFu)
			continue;
		const float *p = predict + 3 * s;
		if ((e & 1u) == 0u) {
			uint32_t xy = e >> 1, px = xy & 0x7FFFu, py = xy >> 15;
			float *bf = bias_factor_r + 4 * ((uint64_t)py * width + px);
			const float *gb = factor_gb + 2 * ((uint64_t)py * width + px);
			float cr = bf[0] + bf[3] * p[0], cg = bf[1] + gb[0] * p[1], cb = bf[2] + gb[1] * p[2];
			bf[0] = cr, bf[1] = cg, bf[2] = cb, bf[3] = 0.0f;
		} else {
			uint32_t v = e >> 1, b = v & 3u, l = (v >> 2) & 0x3FFFu, r = v >> 16;
			for (uint32_t i = l; i <= r; ++i) {
				float *rec = train_records[b] + 10 * (uint64_t)i;
				rec[0] = rec[0] + rec[3] * p[0];
				rec[1] = rec[1] + rec[4] * p[1];
				rec[2] = rec[2] + rec[5] * p[2];
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------------------------------
 * UnpackNRCInput (shader/src/NRCRecord.glsl:98-125) over the scene buffers of shader/src/Scene.glsl:8-71.
 * Pointers are HOST pointers here. Texture fetch = what `texture(sampler2D, uv)` does in a compute shader for an
 * R8G8B8A8_SRGB image with one mip level, VK_FILTER_LINEAR, ADDRESS_MODE_REPEAT: sRGB -> linear per texel, then a
 * bilinear blend of the 2x2 footprint around uv*size - 0.5 (fp32 weights; the hardware quantises them to 8 bits).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { float diffuse[3]; uint32_t diffuse_texture_id; float specular[3]; uint32_t specular_texture_id;
                 float emission[3]; uint32_t emission_texture_id; float metallic, roughness, ior; uint32_t pad; } NrcOracleMaterial;
typedef struct { const void *texels; uint32_t width, height; } NrcOracleTexture;
typedef struct { const float *vertices; const uint32_t *vertex_indices; const float *texcoords; const uint32_t *texcoord_indices;
                 const NrcOracleMaterial *materials; const uint32_t *material_ids; const float *transforms;
                 const NrcOracleTexture *textures; uint32_t texture_count; } NrcOracleScene;

static float srgb_to_linear(uint8_t c) { /* the sRGB EOTF, evaluated in double and rounded once to fp32 */
	double x = (double)c / 255.0;
	return (float)(x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4));
}
static void sample_texture(const NrcOracleTexture *t, float u, float v, float rgb[3]) {
	float x = u * (float)t->width - 0.5f, y = v * (float)t->height - 0.5f;
	float fx = floorf(x), fy = floorf(y), tx = x - fx, ty = y - fy;
	int w = (int)t->width, h = (int)t->height;
	int x0 = (int)fx % w, y0 = (int)fy % h;
	if (x0 < 0) x0 += w;
	if (y0 < 0) y0 += h;
	int x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
	const uint8_t *p = (const uint8_t *)t->texels;
	for (int c = 0; c < 3; ++c) {
		float c00 = srgb_to_linear(p[4 * (y0 * w + x0) + c]), c10 = srgb_to_linear(p[4 * (y0 * w + x1) + c]);
		float c01 = srgb_to_linear(p[4 * (y1 * w + x0) + c]), c11 = srgb_to_linear(p[4 * (y1 * w + x1) + c]);
		rgb[c] = (c00 * (1.0f - tx) + c10 * tx) * (1.0f - ty) + (c01 * (1.0f - tx) + c11 * tx) * ty;
	}
}
static void scene_vertex(const NrcOracleScene *sc, uint32_t instance, uint32_t prim, uint32_t k, float out[3]) { /* Scene.glsl:50-53 */
	const float *v = sc->vertices + 3 * (size_t)sc->vertex_indices[3 * (size_t)prim + k];
	const float *m = sc->transforms + 12 * (size_t)instance; /* 3 columns of vec4; vec4(v,1) * M -> dot with each column */
	for (int j = 0; j < 3; ++j)
		out[j] = v[0] * m[4 * j] + v[1] * m[4 * j + 1] + v[2] * m[4 * j + 2] + m[4 * j + 3];
}
void nrc_oracle_unpack_input(const NrcOracleScene *sc, const uint32_t packed[4], float out14[14]) {
	uint32_t prim = packed[0], instance = packed[1] & 0x7FFFFFFFu;
	int flip = (int)(packed[1] >> 31);
	float v[3][3], tc[3][2];
	for (uint32_t k = 0; k < 3; ++k) {
		scene_vertex(sc, instance, prim, k, v[k]);
		const float *t = sc->texcoords + 2 * (size_t)sc->texcoord_indices[3 * (size_t)prim + k];
		tc[k][0] = t[0], tc[k][1] = t[1];
	}
	float e1[3], e2[3], nrm[3];
	for (int j = 0; j < 3; ++j)
		e1[j] = v[1][j] - v[0][j], e2[j] = v[2][j] - v[0][j];
	nrm[0] = e1[1] * e2[2] - e1[2] * e2[1], nrm[1] = e1[2] * e2[0] - e1[0] * e2[2], nrm[2] = e1[0] * e2[1] - e1[1] * e2[0];
	float inv = 1.0f / sqrtf(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
	for (int j = 0; j < 3; ++j)
		nrm[j] = flip ? -(nrm[j] * inv) : nrm[j] * inv;
	float by = (float)(packed[2] & 0xFFFFu) / 65535.0f, bz = (float)(packed[2] >> 16) / 65535.0f, bx = 1.0f - by - bz; /* :111-112 */
	for (int j = 0; j < 3; ++j)
		out14[j] = v[0][j] * bx + v[1][j] * by + v[2][j] * bz;
	out14[3] = (float)(packed[3] & 0xFFFFu) / 65535.0f, out14[4] = (float)(packed[3] >> 16) / 65535.0f; /* :118 */
	/* NRCSphEncode, :47-49 */
	out14[5] = (nrm[0] == 0.0f && nrm[1] == 0.0f) ? 0.5f : 0.5f + atan2f(nrm[1], nrm[0]) / (2.0f * 3.14159265358979323846f);
	out14[6] = acosf(fminf(fmaxf(nrm[2], -1.0f), 1.0f)) / 3.14159265358979323846f;
	const NrcOracleMaterial *m = sc->materials + sc->material_ids[prim];
	out14[7] = m->roughness;
	float u = tc[0][0] * bx + tc[1][0] * by + tc[2][0] * bz, w = tc[0][1] * bx + tc[1][1] * by + tc[2][1] * bz;
	if (m->diffuse_texture_id == 0xFFFFFFFFu)
		memcpy(out14 + 8, m->diffuse, 12);
	else
		sample_texture(sc->textures + m->diffuse_texture_id, u, w, out14 + 8);
	if (m->specular_texture_id == 0xFFFFFFFFu)
		memcpy(out14 + 11, m->specular, 12);
	else
		sample_texture(sc->textures + m->specular_texture_id, u, w, out14 + 11);
}
void nrc_oracle_unpack_batch(const NrcOracleScene *sc, const uint8_t *packed, uint64_t n, uint32_t stride_bytes, float *out14) {
	for (uint64_t i = 0; i < n; ++i) {
		uint32_t pk[4];
		memcpy(pk, packed + i * stride_bytes, 16);
		nrc_oracle_unpack_input(sc, pk, out14 + 14 * i);
	}
}

int nrc_oracle_weight_count(void) { return NRC_WEIGHTS; }
