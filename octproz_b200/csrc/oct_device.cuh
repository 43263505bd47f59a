/*
 * oct_device.cuh -- sm_100a device helpers: mbarrier + bulk-copy (TMA, UBLKCP) staging, named barriers,
 * exact integer->float conversion on the ALU pipes, kernel argument blocks.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "oct_phases.cuh"

namespace octb200 {

/* ---------------- mbarrier / bulk async copy (cp.async.bulk = TMA 1-D, SASS UBLKCP) ---------------- */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "OCT_WAIT:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra OCT_DONE;\n\t"
	    "bra OCT_WAIT;\n\t"
	    "OCT_DONE:\n\t"
	    "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* global -> shared bulk copy completing on an mbarrier; dst/src 16-byte aligned, bytes % 16 == 0 */
__device__ __forceinline__ void bulk_g2s(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* sub-block barrier among `threads` threads (multiple of 32) */
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

/* ---------------- exact u16 -> fp32 without the conversion pipe ----------------
 * 0x4B000000 | v is the float 2^23 + v; subtracting 2^23 is exact for v < 2^23.  Same value as
 * __uint2float_rd(in[index]) of inputToCufftComplex (cuda_code.cu:118-121). */
__device__ __forceinline__ float u16lo_to_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f; }
__device__ __forceinline__ float u16hi_to_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f; }

/* ---------------- en-face gather over peer memory, fused into the epilogue (multi-GPU shards) ----------------
 * world > 0: the lane that finalises depth bin frameNr of a line (updateDisplayedEnFaceFrame with one frame, cuda_code.cu:909)
 * keeps that output value in a register and writes it into this rank's frame window; at the end of the SAME kernel all CTAs
 * push the finished slab into the frame window of every other rank (coalesced P2P stores over NVLink) and the last one
 * publishes `seq` in every rank's flag word (release, system scope).  See k_aux.cu enface_gather_kernel for the
 * stand-alone form used for multi-frame averages / MIP and when later passes (sinusoidal correction, background recording)
 * still change the slab. */
constexpr int OCT_MAX_PEERS = 16;
struct GatherDev {
	float* frames[OCT_MAX_PEERS];
	unsigned* flags[OCT_MAX_PEERS];
	unsigned* counter;
	unsigned Eglobal, offset, seq;
	unsigned frameNr, nFrames;
	int fn;
	int world, rank;
};
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
/* after ONE __threadfence_system(): a relaxed system-scope store per flag word (fence + relaxed store = release; a st.release per
 * peer would repeat the system fence world times in the tail of the kernel) */
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned* p, unsigned v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

/* phase 1 (per line, by the one lane that holds the depth bin): the en-face value goes into THIS rank's own frame window */
__device__ __forceinline__ void gather_store(const GatherDev& g, unsigned line, float val) {
	g.frames[g.rank][(g.Eglobal - 1u) - (g.offset + line)] = val;
}
/* phase 2 (end of the kernel, all CTAs): once every CTA has finished its lines (grid barrier on g.counter -- the grid is one
 * persistent CTA per SM, all resident), each CTA pushes its share of this rank's slab to every peer with coalesced 16-byte
 * stores over NVLink; the last CTA to finish the push publishes `seq` in every rank's flag word.  Scattered 4-byte peer stores
 * straight from the epilogue cost 34 us per volume at 8 GPUs; the bulk push costs a few. */
__device__ __forceinline__ void gather_push_and_publish(const GatherDev& g, unsigned E) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(g.counter, 1u);
		while (ld_acquire_gpu_u32(g.counter) < gridDim.x) __nanosleep(64);
	}
	__syncthreads();
	if (g.world > 1) {
		const unsigned lo = g.Eglobal - g.offset - E;                       /* this rank's slab inside the (reversed) frame */
		const float* src = g.frames[g.rank] + lo;
		/* head / tail so that the body is 16-byte aligned in every window (windows are 256-byte aligned) */
		const unsigned head = ((4u - (lo & 3u)) & 3u) < E ? ((4u - (lo & 3u)) & 3u) : E;
		const unsigned quads = (E - head) >> 2, tail = head + (quads << 2);
		const unsigned per = (quads + gridDim.x - 1) / gridDim.x;
		const unsigned q0 = blockIdx.x * per, q1 = (q0 + per < quads) ? q0 + per : quads;
		for (int r = 0; r < g.world; ++r) {
			if (r == g.rank) continue;
			float* dst = g.frames[r] + lo;
			for (unsigned q = q0 + threadIdx.x; q < q1; q += blockDim.x)
				reinterpret_cast<float4*>(dst + head)[q] = __ldcg(reinterpret_cast<const float4*>(src + head) + q);
			if (blockIdx.x == 0) {
				for (unsigned i = threadIdx.x; i < head; i += blockDim.x) dst[i] = __ldcg(src + i);
				for (unsigned i = tail + threadIdx.x; i < E; i += blockDim.x) dst[i] = __ldcg(src + i);
			}
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();      /* one system-scope fence after the CTA barrier orders the peer stores of all its threads (fence cumulativity) */
		const unsigned done = atomicAdd(g.counter, 1u);
		if (done == 2u * gridDim.x - 1u) {
			*g.counter = 0;              /* every CTA has left the spin above: safe to rearm for the next launch (stream ordered) */
			__threadfence_system();
#pragma unroll 1
			for (int r = 0; r < g.world; ++r) st_relaxed_sys_u32(g.flags[r] + g.rank, g.seq);
		}
	}
}

/* ---------------- argument block of the fused / own-FFT kernels ---------------- */
struct FusedArgs {
	const uint16_t* raw;     /* SRC_RAW16: [lines][N] u16 (whole raw buffer, halo reads clip to [0,totalSamples)); SRC_RAW12P: [lines][N*3/2] bytes */
	const float2* cin;       /* SRC_CPLX : [lines][N] float2 FFT input written by the pre-FFT kernel */
	float* out;              /* [lines][N/2] processed output slab (flip folded into the line address) */
	float2* cplxOut;         /* != NULL: write the pre-FPN complex bins [lines][N/2] instead (FPN determination pass) */
	const float4* lutB;      /* N entries */
	const float2* tw;        /* 1024 inter-pass twiddles */
	const float2* ctw;       /* 1024 combine twiddles (R == 2) */
	const float2* meanLine;  /* N/2 */
	const float* ppbg;       /* N/2 */
	EpiConsts epi;
	long long totalSamples;
	int lines;
	int A;
	int flip;
	unsigned bscanBase;
	int shiftBits;
	int W;
	int HB, HA;
	GatherDev eg;            /* eg.world == 0: no en-face gather in this launch */
	/* dispersion sweep (gridDim.y = trials): trial t = blockIdx.y processes the SAME lines with its own stage LUT and writes its own output slab */
	int trials;              /* 0 / 1: plain launch */
	int trialLutStride;      /* float4 elements between the LUT images of consecutive trials */
	long long trialOutStride;/* floats between the output slabs of consecutive trials */
	/* floatToOutput (cuda_code.cu:943-967) folded into the epilogue: besides the float line, the same pass writes the line converted to
	 * the acquisition's u16 container (the buffer the stream-to-host D2H copies), so the converted output costs 1 B/sample of extra
	 * writes instead of a second kernel reading the slab back (6 B per output).  NULL: off. */
	unsigned short* convOut; /* [lines][N/2] u16 containers, same line order as `out` */
	float convScale;         /* 2^bits - 1 of cuda_code.cu:948-958 (1023, 4095, 65535) */
};

/* where a line comes from: u16 containers (the reference's raw format), float2 FFT input written by the pre-FFT kernel, or
 * 12-bit samples packed little-endian, two per three bytes (GenICam "Mono12p": an extension -- the reference only takes containers,
 * docs/docs/faq.md -- that cuts the PCIe / HBM input bytes by a quarter) */
enum { SRC_RAW16 = 0, SRC_CPLX = 1, SRC_RAW12P = 2 };
__host__ __device__ constexpr bool src_is_raw(int src) { return src == SRC_RAW16 || src == SRC_RAW12P; }

}  // namespace octb200
