// Micro-benchmark (development only): FFMA vs FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* p, int iters) {
	float2 a[8], b = p[threadIdx.x], c = p[threadIdx.x + 1];
#pragma unroll
	for (int i = 0; i < 8; ++i) a[i] = p[threadIdx.x + 2 + i];
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
			else if (MODE == 1) { a[i] = __ffma2_rn(a[i], b, c); }
			else if (MODE == 2) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }
			else { a[i] = __fadd2_rn(a[i], b); }
		}
	}
	float2 s = a[0];
#pragma unroll
	for (int i = 1; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
	p[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float2* d) {
	const int iters = 4096, blocks = 148 * 8, threads = 256;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k<MODE><<<blocks, threads>>>(d, 16); cudaDeviceSynchronize();
	cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double ops = (double)blocks * threads * iters * 8 * 2;   // scalar fp32 ops (fma counted once)
	printf("%-8s %8.3f ms  %8.2f Tlane-op/s\n", name, ms, ops / ms / 1e9);
}
int main() {
	float2* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2) + 4096); cudaMemset(d, 0, 148 * 8 * 256 * sizeof(float2) + 4096);
	run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FADD", d); run<3>("FADD2", d);
	return 0;
}
