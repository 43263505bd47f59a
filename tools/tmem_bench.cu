// Development micro-benchmark: tensor-memory (TMEM) as a per-lane lookup-table store.
// tcgen05.st fills 256 columns per lane quadrant, then every warp re-reads them with tcgen05.ld 32x32b.xN
// and we measure bytes/clk/SM and check the values.  sm_100a only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t* r) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
	               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
	             :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

template <int X, int DEPTH>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* clocks, int iters, int cols) {
	__shared__ uint32_t tbase_s;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tbase_s)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tbase = tbase_s;
	const uint32_t quad = (uint32_t)(warp & 3) * 32u;
	const uint32_t mybase = tbase + (quad << 16);
	if (warp < 4) {   // each quadrant filled by one warp: value = lane-in-quadrant * 1000 + column
		for (int c = 0; c < cols; c += 8) {
			uint32_t r[8];
			for (int i = 0; i < 8; ++i) r[i] = (quad + lane) * 1000u + (uint32_t)(c + i);
			tmem_st8(mybase + (uint32_t)c, r);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t acc = 0, bad = 0;
	const long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
		for (int c = 0; c < cols; c += X * DEPTH) {
			uint32_t r[DEPTH][X];
#pragma unroll
			for (int d = 0; d < DEPTH; ++d) tmem_ld<X>(mybase + (uint32_t)(c + d * X), r[d]);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for (int d = 0; d < DEPTH; ++d)
#pragma unroll
				for (int i = 0; i < X; ++i) { acc += r[d][i]; if (it == 0 && r[d][i] != (quad + lane) * 1000u + (uint32_t)(c + d * X + i)) bad++; }
		}
	}
	const long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = bad + (acc == 12345u ? 1u : 0u);
	if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

template <int X, int DEPTH> void run(int warps) {
	const int blocks = 148, iters = 200, cols = 256;
	uint32_t* d; long long* c;
	cudaMalloc(&d, blocks * 512 * 4); cudaMalloc(&c, blocks * 8);
	k<X, DEPTH><<<blocks, warps * 32>>>(d, c, iters, cols);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("x%d warps %d: CUDA error %s\n", X, warps, cudaGetErrorString(e)); return; }
	long long hc[148]; cudaMemcpy(hc, c, blocks * 8, cudaMemcpyDeviceToHost);
	static uint32_t ho[148 * 512]; cudaMemcpy(ho, d, blocks * warps * 32 * 4, cudaMemcpyDeviceToHost);
	int bad = 0; for (int i = 0; i < blocks * warps * 32; ++i) if (ho[i] != 0) bad++;
	double bytes = (double)iters * cols * 4 * warps * 32;
	printf("32x32b.x%-2d depth %d warps/SM %2d: %8.1f B/clk/SM  (%lld clk)  wrong-value threads: %d\n", X, DEPTH, warps, bytes / (double)hc[0], hc[0], bad);
	cudaFree(d); cudaFree(c);
}
int main() {
	run<8, 1>(4); run<8, 1>(16); run<8, 2>(16); run<8, 4>(16); run<16, 2>(16); run<8, 4>(4);
	return 0;
}
