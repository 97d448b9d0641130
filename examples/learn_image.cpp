// learn_image.cpp -- the reference's "learning an image" demo (test/mlp_learning_an_image/main.cpp:220-240: per frame one
// TrainPass = clear -> gradient.comp -> optimize.comp, then inference.comp over the 640 x 640 window) driven through the
// C ABI of include/nrc_b200.h. No window: the trained image is written as a PPM and the PSNR against the target printed.
//   usage: learn_image [steps=2000] [image.ppm]        (without a file a procedural test card is learnt)
#include <nrc_b200.h>

#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>

static constexpr uint32_t kWindowSize = 640, kBatch = 16384; // mlp_learning_an_image/main.cpp:25, gradient.comp:47
static constexpr float kLearningRate = 0.01f;                // optimize.comp:26

#define CHECK_NRC(call)                                                                                                \
	do {                                                                                                               \
		if (int rc_ = (call); rc_ != NRC_OK) {                                                                         \
			std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, nrc_last_error());                                \
			return EXIT_FAILURE;                                                                                       \
		}                                                                                                              \
	} while (0)
#define CHECK_CUDA(call)                                                                                               \
	do {                                                                                                               \
		if (cudaError_t e_ = (call); e_ != cudaSuccess) {                                                              \
			std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));                                    \
			return EXIT_FAILURE;                                                                                       \
		}                                                                                                              \
	} while (0)

static bool load_ppm(const std::string &path, std::vector<uint8_t> &rgba, uint32_t &w, uint32_t &h) {
	std::ifstream f(path, std::ios::binary);
	std::string magic;
	int maxv = 0;
	if (!(f >> magic >> w >> h >> maxv) || magic != "P6" || maxv != 255)
		return false;
	f.get();
	std::vector<uint8_t> rgb((size_t)w * h * 3);
	if (!f.read((char *)rgb.data(), (std::streamsize)rgb.size()))
		return false;
	rgba.resize((size_t)w * h * 4);
	for (size_t i = 0; i < (size_t)w * h; ++i)
		rgba[4 * i] = rgb[3 * i], rgba[4 * i + 1] = rgb[3 * i + 1], rgba[4 * i + 2] = rgb[3 * i + 2], rgba[4 * i + 3] = 255;
	return true;
}
static void test_card(std::vector<uint8_t> &rgba, uint32_t &w, uint32_t &h) { // smooth gradients, rings and hard edges
	w = h = 512;
	rgba.resize((size_t)w * h * 4);
	for (uint32_t y = 0; y < h; ++y)
		for (uint32_t x = 0; x < w; ++x) {
			const float u = (x + 0.5f) / w, v = (y + 0.5f) / h, r = std::hypot(u - 0.5f, v - 0.5f);
			const float ring = 0.5f + 0.5f * std::cos(40.0f * r), box = (std::fabs(u - 0.3f) < 0.12f && std::fabs(v - 0.7f) < 0.12f) ? 1.0f : 0.0f;
			uint8_t *p = &rgba[4 * ((size_t)y * w + x)];
			p[0] = (uint8_t)(255.0f * (0.7f * u + 0.3f * ring)), p[1] = (uint8_t)(255.0f * (0.6f * v + 0.4f * box));
			p[2] = (uint8_t)(255.0f * (0.5f * ring + 0.5f * (1.0f - u))), p[3] = 255;
		}
}
// the sampler of mlp_learning_an_image/main.cpp:121-124 (bilinear, clamp to edge) at the centre of window pixel (x, y)
static void sample(const std::vector<uint8_t> &img, uint32_t w, uint32_t h, float u, float v, float rgb[3]) {
	const float x = u * w - 0.5f, y = v * h - 0.5f, fx = std::floor(x), fy = std::floor(y), tx = x - fx, ty = y - fy;
	auto cl = [](int a, int hi) { return a < 0 ? 0 : (a > hi ? hi : a); };
	const int x0 = cl((int)fx, (int)w - 1), x1 = cl((int)fx + 1, (int)w - 1), y0 = cl((int)fy, (int)h - 1), y1 = cl((int)fy + 1, (int)h - 1);
	for (int c = 0; c < 3; ++c) {
		auto px = [&](int xx, int yy) { return (float)img[4 * ((size_t)yy * w + xx) + c]; };
		rgb[c] = ((1 - tx) * (1 - ty) * px(x0, y0) + tx * (1 - ty) * px(x1, y0) + (1 - tx) * ty * px(x0, y1) + tx * ty * px(x1, y1)) / 255.0f;
	}
}

int main(int argc, char **argv) {
	const int steps = argc > 1 ? std::atoi(argv[1]) : 2000;
	std::vector<uint8_t> image;
	uint32_t iw = 0, ih = 0;
	if (argc > 2) {
		if (!load_ppm(argv[2], image, iw, ih)) {
			std::fprintf(stderr, "cannot read %s (binary P6 PPM, maxval 255)\n", argv[2]);
			return EXIT_FAILURE;
		}
	} else
		test_card(image, iw, ih);

	nrc_config_t cfg{kWindowSize, kWindowSize, /*seed: He-normal init, main.cpp:79-88*/ 1234};
	nrc_handle_t nrc = nullptr;
	CHECK_NRC(nrc_create(&cfg, 0, &nrc));
	cudaStream_t stream;
	CHECK_CUDA(cudaStreamCreate(&stream));
	uint8_t *d_image = nullptr, *d_out = nullptr;
	CHECK_CUDA(cudaMalloc(&d_image, image.size()));
	CHECK_CUDA(cudaMalloc(&d_out, (size_t)kWindowSize * kWindowSize * 4));
	CHECK_CUDA(cudaMemcpy(d_image, image.data(), image.size(), cudaMemcpyHostToDevice));

	std::mt19937 rng{42}; // (the reference draws the two push-constant seeds from std::random_device every frame)
	std::vector<uint8_t> out((size_t)kWindowSize * kWindowSize * 4);
	auto psnr = [&]() {
		double se = 0.0;
		for (uint32_t y = 0; y < kWindowSize; ++y)
			for (uint32_t x = 0; x < kWindowSize; ++x) {
				float t[3];
				sample(image, iw, ih, (x + 0.5f) / kWindowSize, (y + 0.5f) / kWindowSize, t);
				for (int c = 0; c < 3; ++c) {
					const double d = out[4 * ((size_t)y * kWindowSize + x) + c] / 255.0 - t[c];
					se += d * d;
				}
			}
		return -10.0 * std::log10(se / (3.0 * kWindowSize * kWindowSize));
	};
	const auto t0 = std::chrono::steady_clock::now();
	for (int s = 1; s <= steps; ++s) { // one frame of the reference's loop: TrainPass, then InferencePass
		CHECK_NRC(nrc_image_train_step(nrc, d_image, iw, ih, rng(), rng(), kBatch, kLearningRate, stream));
		CHECK_NRC(nrc_image_infer(nrc, d_out, kWindowSize, stream));
		if (s == steps || (s & (s - 1)) == 0) {
			CHECK_CUDA(cudaMemcpyAsync(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost, stream));
			CHECK_CUDA(cudaStreamSynchronize(stream));
			std::printf("step %6d  PSNR %.2f dB\n", s, psnr());
		}
	}
	CHECK_CUDA(cudaStreamSynchronize(stream));
	const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::printf("%d frames (train 16384 samples + infer 640x640) in %.3f s = %.0f frames/s (PSNR readbacks included)\n", steps, secs, steps / secs);

	std::ofstream f("learn_image_out.ppm", std::ios::binary);
	f << "P6\n" << kWindowSize << " " << kWindowSize << "\n255\n";
	for (size_t i = 0; i < (size_t)kWindowSize * kWindowSize; ++i)
		f.write((const char *)&out[4 * i], 3);
	cudaFree(d_image), cudaFree(d_out), cudaStreamDestroy(stream);
	nrc_destroy(nrc);
	return EXIT_SUCCESS;
}
