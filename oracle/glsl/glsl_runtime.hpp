// glsl_runtime.hpp -- TEST INFRASTRUCTURE: dispatches a translated compute shader on the CPU. One workgroup = `local_size`
// cooperative fibers (ucontext) on one OS thread, run round-robin in invocation order and switched at barrier(), so the
// shader's shared-memory hand-offs behave as on the GPU; workgroups run one after another (in index order, so that buffer
// atomics are applied in a reproducible order) or, with parallel = true, spread over OpenMP threads.
#pragma once
#include <cstdint>
namespace glsl_rt {
void dispatch(uint32_t num_groups, uint32_t local_size, uint32_t subgroup_size, void (*shader_main)(), bool parallel = false);
}
