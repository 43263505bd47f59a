// Micro-benchmark (development only): does a packed FFMA2 (fma.rn.f32x2) occupy the warp scheduler's dispatch port for one cycle or two?
// Each case interleaves 8 independent FFMA2 chains with K other instructions per FFMA2 (integer ALU, LDS, scalar FFMA, MUFU, MOV-like)
// and reports SMSP cycles per FFMA2.  2.0 with one extra instruction per FFMA2 = the second cycle of the packed op is a free issue slot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_bench tools/issue_bench.cu && tools/issue_bench
#include <cstdio>
#include <cuda_runtime.h>

#define F2(a, b, c) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c))
#define IA(x, y) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define LO(x, y) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(x) : "r"(y))
#define FS(x, y, z) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z))
#define MU(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x))
#define LS(x) asm volatile("ld.shared.b32 %0, [%0];" : "+r"(x))
#define LS4(x, y0, y1, y2) asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%0];" : "+r"(x), "=r"(y0), "=r"(y1), "=r"(y2))

template <int KIND, int K>
__global__ void __launch_bounds__(512, 1) bench(unsigned long long* out, int iters, float seed) {
	__shared__ unsigned sm[4096];
	const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
	/* pointer-chase tables: word i holds the shared address of word (i + 32) (LDS.32 walk, conflict free) resp. of quad (i/4 + 32) */
	for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (KIND == 6) ? sbase + 16 * (((i >> 2) + 32) & 1023) : sbase + 4 * ((i + 32) & 4095);
	__syncthreads();
	unsigned long long a[8], b, c;
	{
		float2 t = make_float2(seed, seed * 0.5f);
		b = *reinterpret_cast<unsigned long long*>(&t);
		c = b;
		for (int i = 0; i < 8; ++i) { float2 u = make_float2(seed + i, seed - i); a[i] = *reinterpret_cast<unsigned long long*>(&u); }
	}
	unsigned x[16];
	float f[16];
	for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x + i; f[i] = seed + i; }
	if (KIND == 2) for (int i = 0; i < 16; ++i) x[i] = sbase + 4 * ((threadIdx.x & 31) + 32 * i);
	if (KIND == 6) for (int i = 0; i < 16; ++i) x[i] = sbase + 16 * ((threadIdx.x & 31) + 32 * i);
	unsigned y0 = 0, y1 = 0, y2 = 0;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			F2(a[i], b, c);
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const int s = (i * K + k) & 15;
				if (KIND == 1) IA(x[s], x[(s + 1) & 15]);
				if (KIND == 2) LS(x[s]);
				if (KIND == 3) FS(f[s], f[(s + 1) & 15], f[(s + 2) & 15]);
				if (KIND == 4) MU(f[s]);
				if (KIND == 5) LO(x[s], x[(s + 1) & 15]);
				if (KIND == 6) LS4(x[s], y0, y1, y2);
			}
		}
	}
	const long long t1 = clock64();
	unsigned long long acc = 0;
	for (int i = 0; i < 8; ++i) acc ^= a[i];
	for (int i = 0; i < 16; ++i) acc ^= x[i] ^ __float_as_uint(f[i]);
	acc ^= y0 ^ y1 ^ y2;
	if (acc == 0x1234567ull) out[1] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}

template <int KIND, int K> void run(const char* name, unsigned long long* d, int threads) {
	const int iters = 2000;
	bench<KIND, K><<<148, threads>>>(d, 10, 1.0f);
	bench<KIND, K><<<148, threads>>>(d, iters, 1.0f);
	unsigned long long h = 0;
	cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
	const double warpsPerSmsp = threads / 128.0;
	/* cycles of one SMSP per FFMA2 issued on it */
	printf("ISSUE_BENCH %-26s warps/SMSP %.0f  K=%d  cycles per FFMA2 (per SMSP) = %.3f\n", name, warpsPerSmsp, K, (double)h / ((double)iters * 8 * warpsPerSmsp));
}

int main() {
	unsigned long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
	for (int threads : {128, 256, 512}) {
		run<0, 0>("FFMA2 alone", d, threads);
		run<1, 1>("FFMA2 + 1 IADD", d, threads);
		run<1, 2>("FFMA2 + 2 IADD", d, threads);
		run<5, 1>("FFMA2 + 1 LOP3", d, threads);
		run<2, 1>("FFMA2 + 1 LDS.32", d, threads);
		run<6, 1>("FFMA2 + 1 LDS.128", d, threads);
		run<3, 1>("FFMA2 + 1 FFMA", d, threads);
		run<3, 2>("FFMA2 + 2 FFMA", d, threads);
		run<4, 1>("FFMA2 + 1 MUFU", d, threads);
	}
	cudaError_t e = cudaDeviceSynchronize();
	printf("ISSUE_BENCH status %s\n", cudaGetErrorString(e));
	return 0;
}
