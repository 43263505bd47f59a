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

/* ---------------- programmatic dependent launch (back-to-back buffers on one stream) ----------------
 * launch_dependents: the next kernel in the stream (launched with the programmatic-serialization attribute) may become resident as soon as
 * SMs free up, i.e. its prologue (tensor-memory allocation, table fill, first line load) overlaps this kernel's tail; it calls
 * grid_dependency_wait before it touches anything the previous kernel writes.  Both are no-ops in a plain launch. */
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

/* ---------------- exact u16 -> fp32 without the conversion pipe ----------------
 * 0x4B000000 | v is the float 2^23 + v; subtracting 2^23 is exact for v < 2^23.  Same value as
 * __uint2float_rd(in[index]) of inputToCufftComplex (cuda_code.cu:118-121). */
__device__ __forceinline__ float u16lo_to_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f; }
__device__ __forceinline__ float u16hi_to_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f; }
/* byte B of a word -> fp32, the same way (u8 containers) */
template <int B> __device__ __forceinline__ float u8_to_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 | B)) - 8388608.0f; }

/* ---------------- en-face gather over peer memory, fused into the epilogue (multi-GPU shards) ----------------
 * world > 0: a line group works through blocks of `lineBlock` CONSECUTIVE lines; after a block, lanes 0 .. lineBlock-1 of its first
 * warp read depth bin frameNr of "their" line (updateDisplayedEnFaceFrame with one frame, cuda_code.cu:909) back from L2 -- the
 * epilogue's stores are ordered before the group's end-of-line barrier -- and store the lineBlock neighbouring en-face values, one
 * coalesced 4*lineBlock-byte store per rank, straight into
 * the frame window of EVERY rank (peer-mapped pointers: P2P stores over NVLink / NVSwitch; the own rank is a plain store).  The
 * stores are spread over the whole kernel, there is no end-of-kernel push and no grid barrier: a CTA that has finished its lines
 * fences once at system scope and bumps a counter; the last one publishes `seq` in every rank's "arrived" word for this rank.
 *
 * Window of a rank (octb200.cu EnfaceGather): [64 words arrived[producer]] [64 words ack[consumer]] ... [frame 0] [frame 1] [frame 2].
 * Frames are used round-robin by sequence number (seq % 3).  Flow control: before a producer stores frame `seq` into a peer's window
 * that peer must have consumed frame seq - 3 (the previous user of the buffer): the consumer's enface_consume_kernel (k_aux.cu, in
 * stream order behind its own rank's producing kernel) waits for all arrived[] words, copies the frame into its private display
 * buffer and THEN writes ack[consumer] = seq into every producer's window; a producer checks its own (local) ack words in the kernel
 * prologue.  Three buffers give a slow rank two steps of slack before it holds the others up.
 * Every spin has a time-out (status word), never a hang.
 * See k_aux.cu enface_gather_kernel for the stand-alone form used for multi-frame averages / MIP and when later passes
 * (sinusoidal correction, background recording) still change the slab. */
constexpr int OCT_MAX_PEERS = 16;
constexpr int OCT_GATHER_ARRIVED = 0, OCT_GATHER_ACK = 64, OCT_GATHER_HEADER_BYTES = 1024;
constexpr int OCT_GATHER_FRAMES = 3;       /* frame buffers per window, used round-robin by sequence number */
constexpr unsigned long long OCT_GATHER_TIMEOUT_NS = 10ull * 1000ull * 1000ull * 1000ull;
struct GatherDev {
	float* frames[OCT_MAX_PEERS];      /* frame buffer seq % OCT_GATHER_FRAMES in the window of every rank */
	unsigned* flags[OCT_MAX_PEERS];    /* header of every rank's window: arrived[] at word 0, ack[] at word 64 */
	unsigned* counter;                 /* local CTA completion counter (zero between launches) */
	unsigned* status;                  /* local: [0] = time-outs waiting for acknowledgements, [1] = time-outs waiting for arrivals */
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
__device__ __forceinline__ unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

/* spin until *word >= want (sequence numbers, wrap-safe) or the time-out; returns false on time-out */
__device__ __forceinline__ bool gather_spin_ge(const unsigned* word, unsigned want) {
	if ((int)(ld_acquire_sys_u32(word) - want) >= 0) return true;
	const unsigned long long t0 = global_timer_ns();
	while ((int)(ld_acquire_sys_u32(word) - want) < 0) {
		__nanosleep(100);
		if (global_timer_ns() - t0 > OCT_GATHER_TIMEOUT_NS) return false;
	}
	return true;
}
/* producer prologue (the lanes of ONE warp per CTA, lane c looks at consumer c: one load latency for all ranks): every consumer has
 * released the frame buffer this launch is about to overwrite */
__device__ __forceinline__ void gather_wait_acks(const GatherDev& g, int lane) {
	if (g.world < 1 || g.seq <= (unsigned)OCT_GATHER_FRAMES) return;       /* (the first three gathers find their buffers unused) */
	if (lane < g.world) {
		const unsigned* acks = g.flags[g.rank] + OCT_GATHER_ACK;
		if (!gather_spin_ge(acks + lane, g.seq - (unsigned)OCT_GATHER_FRAMES)) atomicAdd(g.status, 1u);
	}
}
/* after a block of `cnt` consecutive lines starting at `firstLine`: lanes 0 .. cnt-1 hold their en-face values */
__device__ __forceinline__ void gather_store_block(const GatherDev& g, unsigned firstLine, int cnt, int lane, float val) {
	if (lane < cnt) {
		const unsigned idx = (g.Eglobal - 1u) - (g.offset + firstLine + (unsigned)lane);      /* the reference writes the frame reversed (cuda_code.cu:909) */
#pragma unroll 1
		for (int r = 0; r < g.world; ++r) g.frames[r][idx] = val;
	}
}
/* end of the kernel, after a CTA-wide barrier: one system-scope fence per CTA orders the peer stores of all its threads (fence
 * cumulativity through the barrier); the last CTA to arrive publishes `seq` in every rank's arrived[] word for this rank.  No CTA
 * waits for another one, so nothing here depends on the CTAs being co-resident. */
__device__ __forceinline__ void gather_publish(const GatherDev& g) {
	if (threadIdx.x == 0) {
		__threadfence_system();
		const unsigned done = atomicAdd(g.counter, 1u);
		if (done == gridDim.x * gridDim.y - 1u) {
			*g.counter = 0;              /* the next launch is stream ordered behind this one */
			__threadfence_system();
#pragma unroll 1
			for (int r = 0; r < g.world; ++r) st_relaxed_sys_u32(g.flags[r] + OCT_GATHER_ARRIVED + g.rank, g.seq);
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
	int pdl;                 /* launched with programmatic stream serialization behind another main launch of the same handle */
	int lineBlock;           /* a line group works through blocks of this many consecutive lines (1, 2, 4, 8; 0 = 1); see GatherDev */
	unsigned flipEnd;        /* B-scans with (index + bscanBase) >= flipEnd are never flipped: with an odd number of B-scans the reference's
	                            cuda_bscanFlip leaves the last one alone (cuda_code.cu:794-805 runs over samplesPerBuffer/4 elements) */
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
enum { SRC_RAW16 = 0, SRC_CPLX = 1, SRC_RAW12P = 2, SRC_RAW8 = 3, SRC_RAW32 = 4 };      /* RAW8 / RAW32: u8 / u32 containers (cuda_code.cu:116-125) */
__host__ __device__ constexpr bool src_is_raw(int src) { return src == SRC_RAW16 || src == SRC_RAW12P || src == SRC_RAW8 || src == SRC_RAW32; }
/* bytes of one raw line of n samples in the TMA slot */
__host__ __device__ constexpr int src_line_bytes(int src, int n) { return src == SRC_RAW8 ? n : (src == SRC_RAW32 ? 4 * n : (src == SRC_RAW12P ? n * 3 / 2 : 2 * n)); }

}  // namespace octb200
