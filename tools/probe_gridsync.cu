// probe_gridsync.cu -- development tool: cost of the training kernel's grid barrier on its own (cooperative launch,
// one 288-thread CTA per SM, N barriers back to back; cycles per barrier seen by CTA 0).
#include "../vknrc_b200/csrc/nrc_train.cu"
#include <cstdio>
__global__ void __launch_bounds__(288, 1) k(uint32_t *bar, int n, long long *out) {
	const long long t0 = clock64();
	uint32_t target = 0;
	for (int i = 0; i < n; ++i)
		nrc::grid_sync<false>(bar, target);
	if (blockIdx.x == 0 && threadIdx.x == 0)
		out[0] = (clock64() - t0) / n;
}
int main() {
	uint32_t *bar; long long *out;
	cudaMalloc(&bar, 64); cudaMemset(bar, 0, 64); cudaMalloc(&out, 8);
	for (int grid : {16, 64, 128, 148}) {
		int n = 200;
		cudaMemset(bar, 0, 64);
		void *args[] = {&bar, &n, &out};
		cudaLaunchCooperativeKernel((void *)k, dim3(grid), dim3(288), args, 0, 0);
		long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
		printf("grid %3d: %lld cycles per barrier (%s)\n", grid, h, cudaGetErrorString(cudaGetLastError()));
	}
	return 0;
}
