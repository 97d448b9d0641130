// glsl_runtime.cpp -- TEST INFRASTRUCTURE (see glsl_runtime.hpp / glsl_shim.hpp).
#include "glsl_runtime.hpp"
#include "glsl_shim.hpp"

#include <atomic>
#include <cstdlib>
#include <memory>
#include <ucontext.h>
#include <vector>

thread_local uvec3 gl_LocalInvocationID, gl_GlobalInvocationID, gl_WorkGroupID;
thread_local uint gl_SubgroupID, gl_SubgroupInvocationID;

namespace {
constexpr size_t kStackBytes = 192 * 1024; // the gradient shader keeps 64 cooperative matrices (32 KB) in locals

struct Mat16 {
	float16_t e[256];
};
struct Fiber {
	ucontext_t ctx;
	std::unique_ptr<char[]> stack;
	bool done = false;
	uint32_t mma_index = 0; // cooperative-matrix products issued by this invocation so far
};
struct WorkgroupState {
	ucontext_t scheduler;
	std::vector<Fiber> fibers;
	std::vector<std::vector<Mat16>> memo; // per subgroup: the products its first invocation has computed
	uint32_t current = 0, subgroup_size = 32;
	void (*shader_main)() = nullptr;
	bool use_memo = true;
};
thread_local WorkgroupState *g_wg = nullptr;

void fiber_entry() {
	g_wg->shader_main();
	g_wg->fibers[g_wg->current].done = true;
	swapcontext(&g_wg->fibers[g_wg->current].ctx, &g_wg->scheduler); // never resumed
}

void run_workgroup(WorkgroupState &wg, uint32_t group, uint32_t local_size) {
	g_wg = &wg;
	for (auto &m : wg.memo)
		m.clear();
	for (uint32_t i = 0; i < local_size; ++i) {
		Fiber &f = wg.fibers[i];
		f.done = false, f.mma_index = 0;
		getcontext(&f.ctx);
		f.ctx.uc_stack.ss_sp = f.stack.get(), f.ctx.uc_stack.ss_size = kStackBytes, f.ctx.uc_link = nullptr;
		makecontext(&f.ctx, fiber_entry, 0);
	}
	for (uint32_t alive = local_size; alive;) {
		alive = 0;
		for (uint32_t i = 0; i < local_size; ++i) {
			if (wg.fibers[i].done)
				continue;
			wg.current = i;
			gl_WorkGroupID = uvec3(group, 0, 0), gl_LocalInvocationID = uvec3(i, 0, 0), gl_GlobalInvocationID = uvec3(group * local_size + i, 0, 0);
			gl_SubgroupID = i / wg.subgroup_size, gl_SubgroupInvocationID = i % wg.subgroup_size;
			swapcontext(&wg.scheduler, &wg.fibers[i].ctx);
			alive += wg.fibers[i].done ? 0 : 1;
		}
	}
	g_wg = nullptr;
}
} // namespace

void barrier() {
	WorkgroupState &wg = *g_wg;
	swapcontext(&wg.fibers[wg.current].ctx, &wg.scheduler); // resumed (with the built-ins restored) on the next round
}

// (E1) of glsl_shim.hpp
void glsl_coopmat_muladd_16(const float16_t *a, const float16_t *b, const float16_t *c, float16_t *d) {
	WorkgroupState *wg = g_wg;
	Fiber *f = wg ? &wg->fibers[wg->current] : nullptr;
	if (wg && wg->use_memo) {
		std::vector<Mat16> &memo = wg->memo[wg->current / wg->subgroup_size];
		const uint32_t idx = f->mma_index++;
		if (idx < memo.size()) { // a later invocation of the subgroup: same operands, same product
			std::memcpy(d, memo[idx].e, sizeof(Mat16));
			return;
		}
		memo.emplace_back();
		float bt[16][16];
		for (int k = 0; k < 16; ++k)
			for (int j = 0; j < 16; ++j)
				bt[j][k] = (float)b[k * 16 + j];
		for (int i = 0; i < 16; ++i)
			for (int j = 0; j < 16; ++j) {
				float s = 0.0f;
				for (int k = 0; k < 16; ++k)
					s += (float)a[i * 16 + k] * bt[j][k];
				memo.back().e[i * 16 + j] = (float16_t)((float)c[i * 16 + j] + s);
			}
		std::memcpy(d, memo.back().e, sizeof(Mat16));
		return;
	}
	for (int i = 0; i < 16; ++i)
		for (int j = 0; j < 16; ++j) {
			float s = 0.0f;
			for (int k = 0; k < 16; ++k)
				s += (float)a[i * 16 + k] * (float)b[k * 16 + j];
			d[i * 16 + j] = (float16_t)((float)c[i * 16 + j] + s);
		}
}

void glsl_atomic_add(float *p, float v) { std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }

// (E2) of glsl_shim.hpp
static float srgb_to_linear(uint8_t c) {
	const double x = c / 255.0;
	return (float)(x <= 0.04045 ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4));
}
vec4 texture(const sampler2D &s, vec2 uv) {
	const float x = uv.x * (float)s.width - 0.5f, y = uv.y * (float)s.height - 0.5f;
	const float fx = std::floor(x), fy = std::floor(y), a = x - fx, b = y - fy;
	auto wrap = [&](int i, int n) {
		if (s.repeat) {
			i %= n;
			return i < 0 ? i + n : i;
		}
		return i < 0 ? 0 : (i >= n ? n - 1 : i);
	};
	const int x0 = wrap((int)fx, s.width), x1 = wrap((int)fx + 1, s.width), y0 = wrap((int)fy, s.height), y1 = wrap((int)fy + 1, s.height);
	vec4 r;
	for (int c = 0; c < 4; ++c) {
		auto texel = [&](int xx, int yy) {
			const uint8_t t = s.texels[4 * ((size_t)yy * s.width + xx) + c];
			return (s.srgb && c < 3) ? srgb_to_linear(t) : (float)t / 255.0f;
		};
		r[c] = (1.0f - a) * (1.0f - b) * texel(x0, y0) + a * (1.0f - b) * texel(x1, y0) + (1.0f - a) * b * texel(x0, y1) + a * b * texel(x1, y1);
	}
	return r;
}

namespace glsl_rt {
void dispatch(uint32_t num_groups, uint32_t local_size, uint32_t subgroup_size, void (*shader_main)(), bool parallel) {
	const bool memo = !std::getenv("GLSL_EMU_NO_MEMO");
#pragma omp parallel if (parallel)
	{
		WorkgroupState wg;
		wg.shader_main = shader_main, wg.subgroup_size = subgroup_size, wg.use_memo = memo;
		wg.fibers.resize(local_size);
		for (auto &f : wg.fibers)
			f.stack.reset(new char[kStackBytes]);
		wg.memo.resize((local_size + subgroup_size - 1) / subgroup_size);
#pragma omp for schedule(dynamic, 1)
		for (uint32_t g = 0; g < num_groups; ++g)
			run_workgroup(wg, g, local_size);
	}
}
} // namespace glsl_rt
