// render_loop.cpp -- a renderer-shaped frame loop on the reference's own record formats, through the C++ face of the C ABI
// (include/nrc_b200.hpp, shaped like VkNRCState + the NNInference / NNTrain passes). Per frame, as in
// src/rg/NRCRenderGraph.cpp:46-113: the host zeroes the counts, a producer fills the eval / train record buffers (here:
// random hits on a small procedural scene instead of path_tracer.comp, with an analytic target radiance so that the cache
// has something to learn), then NNInference composites the cache's answer into the screen images and NNTrain runs the four
// batches. Prints the training loss (gradient buffer slot 20672 / 20673) - it must go down - and checks the composite.
#include <nrc_b200.hpp>

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CHECK_CUDA(call)                                                                                               \
	do {                                                                                                               \
		if (cudaError_t e_ = (call); e_ != cudaSuccess) {                                                              \
			std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));                                    \
			return EXIT_FAILURE;                                                                                       \
		}                                                                                                              \
	} while (0)

template <class T> static T *upload(const std::vector<T> &v) {
	T *d = nullptr;
	if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess || cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
		std::abort();
	return d;
}

int main(int argc, char **argv) {
	const int frames = argc > 1 ? std::atoi(argv[1]) : 64;
	const nrc::Extent extent{640, 360};
	const uint32_t n_pix = extent.width * extent.height, kBatch = nrc::State::GetTrainBatchSize(), kBatches = nrc::State::GetTrainBatchCount();
	std::mt19937 rng{7};
	std::uniform_real_distribution<float> uni{0.0f, 1.0f};

	// ---- scene: a 16 x 16 grid of quads in the plane z = 0 (two triangles each), 4 materials, one 64 x 64 checker texture
	const uint32_t grid = 16, n_prims = grid * grid * 2;
	std::vector<float> vertices, texcoords{0, 0, 1, 0, 0, 1, 1, 1};
	std::vector<uint32_t> vidx, tidx, material_ids;
	for (uint32_t y = 0; y <= grid; ++y)
		for (uint32_t x = 0; x <= grid; ++x)
			vertices.insert(vertices.end(), {2.0f * x / grid - 1.0f, 2.0f * y / grid - 1.0f, 0.0f});
	for (uint32_t y = 0; y < grid; ++y)
		for (uint32_t x = 0; x < grid; ++x) {
			const uint32_t a = y * (grid + 1) + x, b = a + 1, c = a + grid + 1, d = c + 1;
			vidx.insert(vidx.end(), {a, b, c, b, d, c});
			tidx.insert(tidx.end(), {0, 1, 2, 1, 3, 2});
			material_ids.insert(material_ids.end(), {(x + y) % 4, (x + y) % 4});
		}
	std::vector<uint32_t> texels(64 * 64);
	for (uint32_t i = 0; i < 64 * 64; ++i)
		texels[i] = (((i % 64) / 8 + (i / 64) / 8) & 1) ? 0xFFE0E0E0u : 0xFF404040u;
	uint32_t *d_texels = upload(texels);
	std::vector<NrcTexture> textures{{d_texels, 64, 64}};
	std::vector<NrcMaterial> materials(4);
	for (uint32_t m = 0; m < 4; ++m) {
		NrcMaterial &mt = materials[m];
		mt = NrcMaterial{};
		mt.diffuse[0] = 0.2f + 0.2f * m, mt.diffuse[1] = 0.8f - 0.2f * m, mt.diffuse[2] = 0.5f;
		mt.specular[0] = mt.specular[1] = mt.specular[2] = 0.04f;
		mt.diffuse_texture_id = m == 3 ? 0u : 0xFFFFFFFFu, mt.specular_texture_id = mt.emission_texture_id = 0xFFFFFFFFu;
		mt.roughness = 0.25f * (m + 1), mt.ior = 1.5f;
	}
	std::vector<float> transforms{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}; // one instance, identity mat3x4
	NrcScene scene{};
	scene.vertices = upload(vertices), scene.vertex_indices = upload(vidx), scene.texcoords = upload(texcoords), scene.texcoord_indices = upload(tidx);
	scene.materials = upload(materials), scene.material_ids = upload(material_ids), scene.transforms = upload(transforms);
	scene.textures = upload(textures), scene.texture_count = 1;
	void *d_table = nullptr;
	CHECK_CUDA(cudaMalloc(&d_table, nrc_scene_prim_table_bytes(n_prims)));

	try {
		nrc::State state(0, extent, /*seed*/ 2024);
		cudaStream_t stream;
		CHECK_CUDA(cudaStreamCreate(&stream));
		nrc::State::PrepareScene(scene, n_prims, d_table, stream);

		// ---- per-frame resources (src/rg/NRCRenderGraph.cpp:139-175)
		nrc::FrameBuffers f;
		void *d_eval = nullptr;
		uint32_t *d_counts = nullptr; // [0] eval count, [1..4] batch counts
		CHECK_CUDA(cudaMalloc(&d_eval, nrc::State::GetEvalRecordBufferSize(extent)));
		CHECK_CUDA(cudaMalloc(&d_counts, 5 * sizeof(uint32_t)));
		CHECK_CUDA(cudaMalloc(&f.bias_factor_r, (size_t)n_pix * 16));
		CHECK_CUDA(cudaMalloc((void **)&f.factor_gb, (size_t)n_pix * 8));
		for (uint32_t b = 0; b < kBatches; ++b) {
			CHECK_CUDA(cudaMalloc(&f.train_records[b], nrc::State::GetBatchTrainRecordBufferSize()));
			f.train_counts[b] = d_counts + 1 + b;
		}
		f.eval_records = d_eval, f.eval_count = d_counts, f.max_eval_count = n_pix + (uint64_t)kBatch * kBatches, f.image_pitch = extent.width;

		auto random_hit = [&]() {
			NrcPackedInput p{};
			p.primitive_id = rng() % n_prims, p.flip_bit_instance_id = 0;
			float by = uni(rng), bz = uni(rng);
			if (by + bz > 1.0f)
				by = 1.0f - by, bz = 1.0f - bz;
			p.barycentric_2x16U = (uint32_t)std::lround(by * 65535.0f) | ((uint32_t)std::lround(bz * 65535.0f) << 16);
			p.scattered_dir_2x16U = (rng() & 0xFFFFu) | ((rng() & 0xFFFFu) << 16);
			return p;
		};
		// the "light transport" the cache has to learn: radiance as a smooth function of where the primitive sits
		auto radiance = [&](const NrcPackedInput &p, float rgb[3]) {
			const uint32_t quad = p.primitive_id / 2, x = quad % grid, y = quad / grid;
			const float cx = (x + 0.5f) / grid, cy = (y + 0.5f) / grid;
			rgb[0] = 0.5f + 0.5f * std::sin(6.0f * cx), rgb[1] = 0.5f + 0.5f * std::cos(5.0f * cy), rgb[2] = cx * cy;
		};
		std::vector<NrcEvalRecord> eval(n_pix);
		std::vector<NrcTrainRecord> train(kBatch);
		std::vector<float> image((size_t)n_pix * 4), gradients(NRC_B200_GRADIENT_FLOATS);
		float first_loss = 0.0f, last_loss = 0.0f;
		double composite_err = 0.0;
		for (int frame = 0; frame < frames; ++frame) {
			state.NextFrame();
			// producer: a screen query per pixel (dst = (x | y << 15) << 1, NRCRecord.glsl:19-24) and four full train batches
			for (uint32_t i = 0; i < n_pix; ++i)
				eval[i] = NrcEvalRecord{((i % extent.width) | ((i / extent.width) << 15)) << 1, random_hit()};
			CHECK_CUDA(cudaMemcpyAsync(d_eval, eval.data(), eval.size() * sizeof(NrcEvalRecord), cudaMemcpyHostToDevice, stream));
			const uint32_t counts[5] = {n_pix, kBatch, kBatch, kBatch, kBatch};
			CHECK_CUDA(cudaMemcpyAsync(d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice, stream));
			for (uint32_t b = 0; b < kBatches; ++b) {
				for (auto &t : train) {
					t.packed_input = random_hit();
					float rgb[3];
					radiance(t.packed_input, rgb);
					t.bias_r = rgb[0], t.bias_g = rgb[1], t.bias_b = rgb[2], t.factor_r = t.factor_g = t.factor_b = 0.0f;
				}
				CHECK_CUDA(cudaMemcpyAsync(f.train_records[b], train.data(), train.size() * sizeof(NrcTrainRecord), cudaMemcpyHostToDevice, stream));
				CHECK_CUDA(cudaStreamSynchronize(stream)); // (`train` is reused by the next batch)
			}
			// screen images: bias = 0, factor = 1 -> the composite is the cache's prediction itself (nrc_inference.comp:53-59)
			for (uint32_t i = 0; i < n_pix; ++i)
				image[4 * i] = image[4 * i + 1] = image[4 * i + 2] = 0.0f, image[4 * i + 3] = 1.0f;
			CHECK_CUDA(cudaMemcpyAsync(f.bias_factor_r, image.data(), image.size() * 4, cudaMemcpyHostToDevice, stream));
			std::vector<float> ones((size_t)n_pix * 2, 1.0f);
			CHECK_CUDA(cudaMemcpyAsync((void *)f.factor_gb, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, stream));

			state.Frame(f, scene, stream); // NNInference, then the four NNTrain pass groups

			state.Download(nullptr, nullptr, nullptr, nullptr, gradients.data(), stream); // synchronises
			const float loss = gradients[NRC_B200_GRAD_LOSS_SLOT] / gradients[NRC_B200_GRAD_COUNT_SLOT];
			if (frame == 0)
				first_loss = loss;
			last_loss = loss;
			if ((frame & (frame + 1)) == 0 || frame + 1 == frames)
				std::printf("frame %4d  relative-L2 loss of the last batch %.5f\n", frame, loss);
			if (frame + 1 == frames) { // how close is the composited image to the radiance the records were trained on?
				CHECK_CUDA(cudaMemcpy(image.data(), f.bias_factor_r, image.size() * 4, cudaMemcpyDeviceToHost));
				for (uint32_t i = 0; i < n_pix; ++i) {
					float rgb[3];
					radiance(eval[i].packed_input, rgb);
					for (int c = 0; c < 3; ++c)
						composite_err += std::fabs(image[4 * i + c] - rgb[c]);
				}
				composite_err /= 3.0 * n_pix;
			}
		}
		std::printf("loss %.5f -> %.5f, mean |composite - radiance| = %.4f\n", first_loss, last_loss, composite_err);
		const bool ok = last_loss < 0.5f * first_loss && std::isfinite(composite_err) && composite_err < 0.25;
		std::printf("%s\n", ok ? "OK" : "FAILED");
		return ok ? EXIT_SUCCESS : EXIT_FAILURE;
	} catch (const nrc::Error &e) {
		std::fprintf(stderr, "nrc error %d: %s\n", e.code, e.what());
		return EXIT_FAILURE;
	}
}
