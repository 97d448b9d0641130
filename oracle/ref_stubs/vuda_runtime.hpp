// Build-only stand-in for the reference's `vuda_runtime.hpp` (test/vuda, a CUDA-shaped API over Vulkan that
// needs Vulkan-Hpp and a Vulkan loader, neither of which exists in this image).
//
// TEST INFRASTRUCTURE. It lets `oracle/Makefile` compile /root/reference/test/main.cpp *where it lies* so that
// the reference's own CPU functions `Evaluate` (test/main.cpp:11-27) and `Train` (test/main.cpp:29-74) can be
// called from oracle/ref_shim.cpp. The GPU half of that file (test_inference/test_train, main.cpp:91-228) only
// has to compile; it is never called, so every entry point below aborts if reached.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <span>
#include <string>
#include <vector>

enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
[[noreturn]] inline void vknrc_ref_stub_unreachable(const char *what) {
	std::fprintf(stderr, "oracle/_ref: %s reached, but the reference GPU path cannot run here\n", what);
	std::abort();
}
inline int cudaMalloc(void **, std::size_t) { vknrc_ref_stub_unreachable("cudaMalloc"); }
inline int cudaMemcpy(void *, const void *, std::size_t, cudaMemcpyKind) { vknrc_ref_stub_unreachable("cudaMemcpy"); }
inline int cudaStreamSynchronize(int) { vknrc_ref_stub_unreachable("cudaStreamSynchronize"); }
inline int cudaSetDevice(int) { vknrc_ref_stub_unreachable("cudaSetDevice"); }
namespace vuda {
template <typename... Args> inline void launchKernel(Args &&...) { vknrc_ref_stub_unreachable("vuda::launchKernel"); }
} // namespace vuda
