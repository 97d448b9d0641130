// C-ABI shim over the reference's own CPU MLP (TEST INFRASTRUCTURE; built only into oracle/_ref/).
//
// `Evaluate` and `Train` are defined in /root/reference/test/main.cpp:11-27 and :29-74; that translation unit is
// compiled unmodified by oracle/Makefile (with -Dmain=vknrc_ref_main and oracle/ref_stubs on the include path) and
// linked with this file. Nothing here re-implements the algorithm: the two functions below only marshal pointers.
#include <cstdint>
#include <cstring>
#include <span>
#include <vector>

using half = _Float16;
std::vector<half> Evaluate(std::span<half> weights, std::span<half> inputs);                         // test/main.cpp:11
std::vector<float> Train(std::span<half> weights, std::span<half> inputs, std::span<half> targets); // test/main.cpp:29

extern "C" {
// weights: 20672 fp16 (row-major W[l][out][in]); inputs: n*64 fp16 sample-major; outputs: n*3 fp16 sample-major.
int vknrc_ref_evaluate(const uint16_t *weights, const uint16_t *inputs, uint64_t n, uint16_t *outputs) {
	std::vector<half> w(20672), x(n * 64);
	std::memcpy(w.data(), weights, w.size() * 2);
	std::memcpy(x.data(), inputs, x.size() * 2);
	std::vector<half> y = Evaluate(w, x);
	if (y.size() != n * 3)
		return -1;
	std::memcpy(outputs, y.data(), y.size() * 2);
	return 0;
}
// dw: 20672 fp32. NOTE: the reference's CPU `Train` is a debugging sketch, not a faithful backward
// pass; it is exposed so it can be *timed* as the reference's CPU cost and so its layer-5 dW can be cross-checked.
int vknrc_ref_train(const uint16_t *weights, const uint16_t *inputs, const uint16_t *targets, uint64_t n, float *dw) {
	std::vector<half> w(20672), x(n * 64), t(n * 3);
	std::memcpy(w.data(), weights, w.size() * 2);
	std::memcpy(x.data(), inputs, x.size() * 2);
	std::memcpy(t.data(), targets, t.size() * 2);
	std::vector<float> g = Train(w, x, t);
	if (g.size() != 20672)
		return -1;
	std::memcpy(dw, g.data(), g.size() * 4);
	return 0;
}
}
