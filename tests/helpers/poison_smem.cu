// Test helper (not product code): fills the whole dynamic shared memory of every SM with the fp16 NaN pattern 0x7FFF, so that
// a kernel relying on "whatever was there" instead of its own initialisation reads poison (tests/test_gpu_frame.py).
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void poison_kernel(uint32_t bytes, uint32_t *sink) {
	extern __shared__ uint32_t sm[];
	for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x)
		sm[i] = 0x7FFF7FFFu;
	__syncthreads();
	if (sm[(threadIdx.x * 97u) % (bytes / 4)] != 0x7FFF7FFFu) // (keeps the stores alive)
		*sink = 1;
}
extern "C" int poison_shared_memory(void *stream) {
	int dev = 0, sms = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint32_t bytes = 227 * 1024;
	static uint32_t *sink = nullptr;
	if (!sink)
		cudaMalloc(&sink, 4);
	cudaFuncSetAttribute(poison_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	poison_kernel<<<sms, 256, bytes, (cudaStream_t)stream>>>(bytes, sink);
	return (int)cudaGetLastError();
}
