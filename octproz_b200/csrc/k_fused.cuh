/*
 * k_fused.cuh -- the hot kernel: one launch takes raw u16 spectra to finished B-scan lines.
 *
 *   raw line --cp.async.bulk (TMA 1-D) + mbarrier, one prefetched slot per line group--> shared memory
 *     -> exact u16->fp32 slot conversion [+ rolling-mean background removal]          (cuda_code.cu:109-211)
 *     -> 4-tap / 16-tap resampling x window x dispersion phasor from LUTs             (cuda_code.cu:213-489)
 *     -> 32x32 four-step inverse FFT, registers + one shared-memory transpose          (cuda_code.cu:1514-1515)
 *     -> fixed-pattern-noise subtract, |.|^2, log/linear scale, truncate to N/2,
 *        B-scan flip folded into the store address, optional background removal        (cuda_code.cu:567-584,699-807)
 *     -> coalesced fp32 stores.   HBM traffic: 2 B in + 2 B out per raw sample, nothing in between.
 *
 * A "line group" is R = N/1024 warps working on one A-scan (R = 1: N = 1024, R = 2: N = 2048: two interleaved
 * 1024-point transforms combined by one radix-2 step).  Groups never synchronise with each other: each owns
 * its two TMA slots, its mbarriers and its exchange tile; the CTA is persistent and strides over the lines.
 *
 * SRC_CPLX variant (OCTB200_FFT_SPLIT): the FFT input comes as float2 from HBM (written by the pre-FFT
 * kernel) -- "own Stockham-style FFT with fused epilogue", BASELINE config 3.
 */
#pragma once
#include "oct_device.cuh"
#include "oct_tmem.cuh"

namespace octb200 {

/* OCT_TMEM_LUT = 1: per-lane tables (stage LUT, twiddles, FPN line, background) live in tensor memory (oct_tmem.cuh)
 * instead of shared memory; 0 keeps everything in shared memory (A/B builds, and the layout the CPU emulator models). */
#ifndef OCT_TMEM_LUT
#define OCT_TMEM_LUT 1
#endif
#ifndef OCT_RT_HALO
#define OCT_RT_HALO 0
#endif
/* experiments for the two-warp kernels (tools/variant_sweep.sh): each of the two instruction cuts of the N = 1024 kernels on its own */
#ifndef OCT_R2_NOSHIFT
#define OCT_R2_NOSHIFT 1
#endif
#ifndef OCT_R2_EGVAR
#define OCT_R2_EGVAR 0
#endif
#ifndef OCT_CVT_I2F
#define OCT_CVT_I2F 0
#endif
#ifndef OCT_R1_THREADS
#define OCT_R1_THREADS 512
#endif
#ifndef OCT_R2_THREADS
#define OCT_R2_THREADS 512
#endif
/* threads per CTA (one CTA per SM): 16 warps x 128 registers fill the register file.  R = 2 used to run 12 warps at 168 registers;
 * after the table reads moved to tensor memory its cubic / linear / plain variants need 124, so they run 16 warps as well
 * (8 line groups instead of 6).  The 16-tap Lanczos stage keeps 12 warps: at 128 registers it spills. */
#ifndef OCT_R2_THREADS_LANCZOS
#define OCT_R2_THREADS_LANCZOS 384
#endif
constexpr int fused_max_threads(int R, int sa) {
	return (R == 1) ? OCT_R1_THREADS : (sa == 2 /* SA_LANCZOS */ ? OCT_R2_THREADS_LANCZOS : OCT_R2_THREADS);
}
template <int R> struct FusedCfg {
	static constexpr int N = 1024 * R;
};

/* shared-memory layout, shared by host (sizing) and device (carving) */
struct FusedSmem {
	int offW, offB, offTw, offCtw, offMean, offPpbg, offGroups;
	int slotBytes, workBytes, groupBytes;
	int total;
};
__host__ __device__ inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

__host__ __device__ inline FusedSmem fused_smem_layout(int R, int sa, bool roll, int src, int HB, int HA, int groups) {
	const int N = 1024 * R, H = N / 2;
	FusedSmem L;
	int off = 0;
	L.offW = off;
#if OCT_TMEM_LUT
	L.offB = off; L.offTw = off; L.offCtw = off; L.offMean = off; L.offPpbg = off;
	off += 128;      /* TMEM base address word */
	(void)H;
#else
	L.offB = off;    off += (src_is_raw(src)) ? 2 * N * 16 : 0;
	L.offTw = off;   off += 1024 * 8;
	L.offCtw = off;  off += (R == 2) ? 1024 * 8 : 0;
	L.offMean = off; off += H * 8;
	L.offPpbg = off; off += H * 4;
#endif
	off = align_up(off, 128);
	L.offGroups = off;
	const int SE = HB + N + HA;
	L.slotBytes = (src_is_raw(src)) ? align_up(src_line_bytes(src, SE), 128) : 0;
	int work = R * XBUF_BYTES;
	if (src_is_raw(src)) {
		int w2 = align_up((FSLOT_PAD + SE + (R == 2 ? 4 : 0)) * 4, 16) + (roll ? align_up((SE + 1) * 4, 16) : 0);   /* R = 2: head room of the odd half */
		if (w2 > work) work = w2;
	}
	L.workBytes = align_up(work, 128);
	L.groupBytes = L.slotBytes + L.workBytes + 128 /* mbarrier */;
	L.total = L.offGroups + groups * L.groupBytes;
	return L;
}

template <int R>
__device__ __forceinline__ void group_sync(int barId) {
	if constexpr (R == 1) __syncwarp();
	else named_bar_sync(barId, 32 * R);
}

/* one lane: start the asynchronous load of raw line `gline` (with halos) into `slot` */
template <int R, bool HALO, int SRC = SRC_RAW16>
__device__ __forceinline__ void issue_line_load(const FusedArgs& a, int gline, unsigned char* slot, uint64_t* bar) {
	constexpr int N = 1024 * R;
	if constexpr (!HALO) {
		/* no halo (every stage but Lanczos): one aligned copy of exactly the line, nothing to clip */
		constexpr uint32_t LB = (uint32_t)src_line_bytes(SRC, N);
		mbar_arrive_expect_tx(bar, LB);
		bulk_g2s(slot, reinterpret_cast<const unsigned char*>(a.raw) + (size_t)gline * LB, LB, bar);
		return;
	}
	const long long lo = (long long)gline * N - a.HB;
	const long long hi = (long long)gline * N + N + a.HA;
	const long long clo = lo < 0 ? 0 : lo;
	const long long chi = hi > a.totalSamples ? a.totalSamples : hi;
	if (clo > lo) { for (long long q = 0; q < clo - lo; q += 8) *reinterpret_cast<uint4*>(slot + q * 2) = make_uint4(0, 0, 0, 0); }
	if (chi < hi) { for (long long q = chi - lo; q < hi - lo; q += 8) *reinterpret_cast<uint4*>(slot + q * 2) = make_uint4(0, 0, 0, 0); }
	const uint32_t bytes = (uint32_t)((chi - clo) * 2);
	mbar_arrive_expect_tx(bar, bytes);
	bulk_g2s(slot + (clo - lo) * 2, a.raw + clo, bytes, bar);
}

/* CONV: the epilogue also writes the line converted to u16 containers (floatToOutput, cuda_code.cu:943-967) into a.convOut.  A kernel
 * template parameter, not a run-time branch: the kernels without conversion stay instruction-for-instruction what they were. */
template <int R, int SA, bool ROLL, int SRC, bool CONV = false>
__global__ void __launch_bounds__(fused_max_threads(R, SA), 1) oct_fused_kernel(const FusedArgs a) {
	constexpr int N = 1024 * R;
	constexpr int H = N / 2;
	/* halos exist only for the 16-tap Lanczos stage: compile-time zero otherwise */
	const int HBv = (SA == SA_LANCZOS || OCT_RT_HALO) ? a.HB : 0, HAv = (SA == SA_LANCZOS || OCT_RT_HALO) ? a.HA : 0;
	extern __shared__ __align__(128) unsigned char smem[];

	/* the warp index through a shuffle: tells the compiler it is warp-uniform, so the line bookkeeping (group, line number,
	 * slot / barrier addresses, loop counter) lives in uniform registers instead of competing with the 64 FFT registers */
	const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
	const int groupsPerCta = (blockDim.x >> 5) / R;
	const int grp = warp / R;          /* line group within the CTA */
	const int p = warp % R;            /* warp within the group = sub-sequence parity */
	const int tig = p * 32 + lane;     /* thread in group */
	const FusedSmem L = fused_smem_layout(R, SA, ROLL, SRC, HBv, HAv, groupsPerCta);

	const float4* sB = reinterpret_cast<const float4*>(smem + L.offB);
	const float2* sTw = reinterpret_cast<const float2*>(smem + L.offTw);
	const float2* sCtw = reinterpret_cast<const float2*>(smem + L.offCtw);
	const float2* sMean = reinterpret_cast<const float2*>(smem + L.offMean);
	const float* sPpbg = reinterpret_cast<const float*>(smem + L.offPpbg);

	/* ---- one-time table fill ---- */
#if OCT_TMEM_LUT
	uint32_t* tmemBaseSlot = reinterpret_cast<uint32_t*>(smem + L.offW);
	if (warp == 0) tmem_alloc(tmemBaseSlot, TmemMap<R>::ALLOC);
	/* multi-GPU en-face gather: every consumer has released the frame buffer this launch overwrites (flow control, oct_device.cuh) */
	if (a.eg.world > 0 && warp == (int)(blockDim.x >> 5) - 1) gather_wait_acks(a.eg, lane);
	tmem_fence_before_sync();
	__syncthreads();
	tmem_fence_after_sync();
	const uint32_t tmemBase = *tmemBaseSlot;
	const uint32_t tq = tmemBase + ((uint32_t)(warp & 3) << 21);     /* lane quadrant of this warp: lane field = bits 31:16, 32 lanes per quadrant */
	{
		const int nWarps = blockDim.x >> 5, quadrant = warp & 3;
		tmem_fill<R>(tq, quadrant, warp >> 2, (nWarps - quadrant + 3) >> 2, lane, a, a.lutB + (size_t)blockIdx.y * a.trialLutStride, src_is_raw(SRC));
	}
	tmem_fence_before_sync();
	__syncthreads();
	tmem_fence_after_sync();
#else
	{
		auto fill = [&](int off, const void* src, int bytes) {
			const uint4* s = reinterpret_cast<const uint4*>(src);
			uint4* d = reinterpret_cast<uint4*>(smem + off);
			for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = __ldg(s + i);
		};
		if constexpr (src_is_raw(SRC)) fill(L.offB, a.lutB + (size_t)blockIdx.y * a.trialLutStride, 2 * N * 16);
		fill(L.offTw, a.tw, 1024 * 8);
		if constexpr (R == 2) fill(L.offCtw, a.ctw, 1024 * 8);
		if (a.epi.fpn && a.cplxOut == nullptr) fill(L.offMean, a.meanLine, H * 8);
		if (a.epi.ppbg) fill(L.offPpbg, a.ppbg, H * 4);
	}
#endif

	unsigned char* gbase = smem + L.offGroups + grp * L.groupBytes;
	/* ONE raw slot per group: it is free again as soon as the line has been converted to fp32, i.e. a full line time
	 * (~10 us with 16 groups per SM) before its next use -- far longer than the HBM latency of the bulk copy */
	unsigned char* slot = gbase;
	unsigned char* work = gbase + L.slotBytes;
	uint64_t* bar = reinterpret_cast<uint64_t*>(work + L.workBytes);
	float2* tile = reinterpret_cast<float2*>(work) + p * XBUF_FLOAT2;
	float2* partnerTile = reinterpret_cast<float2*>(work) + (R == 2 ? (1 - p) : 0) * XBUF_FLOAT2;
	const int barId = 1 + grp;

	const int SE = HBv + N + HAv;
	float* fslot = reinterpret_cast<float*>(work) + FSLOT_PAD;       /* slot element 0; FSLOT_PAD floats of head room */
	unsigned* prefix = reinterpret_cast<unsigned*>(work + align_up((FSLOT_PAD + SE + (R == 2 ? 4 : 0)) * 4, 16));
	/* R = 2, 4-tap stage, no rolling mean: the float slot is split by sample parity (see sample_taps4_x2) */
	constexpr bool SPLIT = (src_is_raw(SRC)) && stage_a_splits_slot(R, SA, ROLL);

	/* line schedule: group g of G works through blocks of LB consecutive lines, block q = g, g + G, g + 2G, ...  LB = 1 is the plain
	 * strided walk; a larger block lets the en-face gather store LB neighbouring values at once (oct_device.cuh GatherDev) */
	const int G = gridDim.x * groupsPerCta;
	const int LB = a.lineBlock > 1 ? a.lineBlock : 1;
	const int blockStep = (G - 1) * LB + 1;            /* from the last line of a block to the first line of the group's next block */
	const int g0 = (blockIdx.x * groupsPerCta + grp) * LB;

	if constexpr (src_is_raw(SRC)) {
		if (tig == 0) { mbar_init(bar, 1); mbar_fence_init(); }
	}
	__syncthreads();
	if constexpr (src_is_raw(SRC)) {
		if (tig == 0 && g0 < a.lines) issue_line_load<R, SA == SA_LANCZOS, SRC>(a, g0, slot, bar);
	}
	/* back-to-back buffers: everything above only read tables and the raw input; the previous kernel's output slab, converted lines and
	 * gather frames are written below.  (No-ops in a plain launch.) */
	grid_launch_dependents();
	grid_dependency_wait();

#ifdef OCT_STAGGER_NS
	/* de-phase the line groups of a CTA so that shared-memory-heavy (stage A) and FMA-heavy (FFT) phases of different
	 * groups overlap instead of all groups hitting the same pipe at the same time */
	__nanosleep((unsigned)(grp * OCT_STAGGER_NS));
#endif
	/* en-face gather fused into the epilogue: k2 index of the displayed depth bin (bin = lane + 32 k2), -1 = off */
	const int egK2 = (a.eg.world > 0) ? (int)(a.eg.frameNr >> 5) : -1;
	int it = 0, jb = 0;                /* jb: position inside the current block */
	for (int gline = g0; gline < a.lines; ++it) {
		const int glineNext = gline + ((jb + 1 == LB) ? blockStep : 1);
		float2 v[32];

		if constexpr (src_is_raw(SRC)) {
			mbar_wait(bar, (uint32_t)(it & 1));

			/* ---- slot conversion: inputToCufftComplex[_and_bitshift] once per sample (cuda_code.cu:109-147) ---- */
			if constexpr (SRC == SRC_RAW12P) {
				/* 12-bit packed, little endian: eight samples in three 32-bit words; 4 octets per thread and 1024 samples */
				static_assert(SA != SA_LANCZOS && !ROLL, "packed input: the host unpacks for the halo / rolling-mean stages");
				const unsigned* s1 = reinterpret_cast<const unsigned*>(slot);
				float4* f4 = reinterpret_cast<float4*>(fslot);
				unsigned w[4][3];
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int o = tig + 32 * R * i;
					w[i][0] = s1[3 * o]; w[i][1] = s1[3 * o + 1]; w[i][2] = s1[3 * o + 2];
				}
				auto f12 = [](unsigned v) { return __uint_as_float(0x4B000000u | (v & 0xFFFu)) - 8388608.0f; };
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					const int o = tig + 32 * R * i;
					const unsigned a0 = w[i][0], a1 = w[i][1], a2 = w[i][2];
					const float x0 = f12(a0), x1 = f12(a0 >> 12), x2 = f12(__funnelshift_r(a0, a1, 24)), x3 = f12(a1 >> 4);
					const float x4 = f12(a1 >> 16), x5 = f12(__funnelshift_r(a1, a2, 28)), x6 = f12(a2 >> 8), x7 = f12(a2 >> 20);
					if constexpr (SPLIT) {
						float4* e4 = reinterpret_cast<float4*>(fslot);
						float4* o4 = reinterpret_cast<float4*>(fslot + SPLIT_ODD_BASE);
						e4[o] = make_float4(x0, x2, x4, x6);
						o4[o] = make_float4(x1, x3, x5, x7);
					} else {
						f4[2 * o] = make_float4(x0, x1, x2, x3);
						f4[2 * o + 1] = make_float4(x4, x5, x6, x7);
					}
				}
				if constexpr (SA == SA_CUBIC) {
					/* mirrored first tap of the cubic: f[-1] = f[1] (cuda_code.cu:284) */
					if (tig == 0) fslot[SPLIT ? SPLIT_ODD_BASE - 1 : -1] = f12(s1[0] >> 12);
				}
			} else if constexpr (SRC == SRC_RAW8 || SRC == SRC_RAW32) {
				/* u8 / u32 containers (cuda_code.cu:116-125,136-146): four samples per thread and step, the same float-slot layout as u16 */
				static_assert(SA != SA_LANCZOS && !ROLL, "u8 / u32 containers: the host takes the split chain for the halo / rolling-mean stages");
				auto put = [&](int q, float4 c) {
					if constexpr (SPLIT) {
						reinterpret_cast<float2*>(fslot)[q] = make_float2(c.x, c.z);
						reinterpret_cast<float2*>(fslot + SPLIT_ODD_BASE)[q] = make_float2(c.y, c.w);
					} else {
						reinterpret_cast<float4*>(fslot)[q] = c;
					}
				};
				const unsigned sh = (unsigned)a.shiftBits;
				if constexpr (SRC == SRC_RAW8) {
					const unsigned* s1 = reinterpret_cast<const unsigned*>(slot);
					unsigned w8[8];
#pragma unroll
					for (int i = 0; i < 8; ++i) w8[i] = s1[tig + 32 * R * i];
					const unsigned msk = (0xFFu >> sh) * 0x01010101u;
#pragma unroll
					for (int i = 0; i < 8; ++i) {
						const unsigned w = (w8[i] >> sh) & msk;
						put(tig + 32 * R * i, make_float4(u8_to_float<0>(w), u8_to_float<1>(w), u8_to_float<2>(w), u8_to_float<3>(w)));
					}
					if constexpr (SA == SA_CUBIC) { if (tig == 0) fslot[SPLIT ? SPLIT_ODD_BASE - 1 : -1] = (float)(slot[1] >> sh); }
				} else {
					const uint4* s4 = reinterpret_cast<const uint4*>(slot);
					auto c32 = [&](unsigned v) { return sh ? (float)((double)v / 4294967296.0) : __uint2float_rd(v); };      /* cuda_code.cu:124,144 */
#pragma unroll 2
					for (int i = 0; i < 8; ++i) {
						const uint4 w = s4[tig + 32 * R * i];
						put(tig + 32 * R * i, make_float4(c32(w.x), c32(w.y), c32(w.z), c32(w.w)));
					}
					if constexpr (SA == SA_CUBIC) { if (tig == 0) fslot[SPLIT ? SPLIT_ODD_BASE - 1 : -1] = c32(reinterpret_cast<const unsigned*>(slot)[1]); }
				}
			} else {
				const uint2* s2 = reinterpret_cast<const uint2*>(slot);
				float4* f4 = reinterpret_cast<float4*>(fslot);
				const unsigned sh = (unsigned)a.shiftBits;
				const unsigned msk = (0xFFFFu >> sh) * 0x00010001u;
				auto cvt = [&](uint2 w) {
					const unsigned wx = (w.x >> sh) & msk, wy = (w.y >> sh) & msk;
					return make_float4(u16lo_to_float(wx), u16hi_to_float(wx), u16lo_to_float(wy), u16hi_to_float(wy));
				};
				/* without the bitshift (the usual case) the shift and mask are identities: one uniform branch per line saves two ALU
				 * instructions per 32-bit word (32 of the ~1230 per line) */
#if OCT_CVT_I2F
				/* experiment: one conversion-pipe instruction per sample (I2F.U16 on a register half) instead of PRMT + FADD */
				auto cvt0 = [](uint2 w) {
					return make_float4((float)(unsigned short)(w.x & 0xFFFFu), (float)(unsigned short)(w.x >> 16),
					                   (float)(unsigned short)(w.y & 0xFFFFu), (float)(unsigned short)(w.y >> 16));
				};
#else
				auto cvt0 = [](uint2 w) { return make_float4(u16lo_to_float(w.x), u16hi_to_float(w.x), u16lo_to_float(w.y), u16hi_to_float(w.y)); };
#endif
				/* first N samples of the slot: 8 quads per thread, all loads in flight before the first conversion */
				uint2 w8[8];
#pragma unroll
				for (int i = 0; i < 8; ++i) w8[i] = s2[tig + 32 * R * i];
				auto store8 = [&](auto conv) {
					if constexpr (SPLIT) {
						/* samples 4q .. 4q+3 -> E[2q], E[2q+1] and O[2q], O[2q+1] */
						float2* e2 = reinterpret_cast<float2*>(fslot);
						float2* o2 = reinterpret_cast<float2*>(fslot + SPLIT_ODD_BASE);
#pragma unroll
						for (int i = 0; i < 8; ++i) {
							const float4 c = conv(w8[i]);
							e2[tig + 32 * R * i] = make_float2(c.x, c.z);
							o2[tig + 32 * R * i] = make_float2(c.y, c.w);
						}
					} else {
#pragma unroll
						for (int i = 0; i < 8; ++i) f4[tig + 32 * R * i] = conv(w8[i]);
					}
				};
				/* N = 1024: 0.2153 -> 0.2112 ms (round 1).  N = 2048: within the run-to-run noise of the two-warp kernel (round 1: 0.5042 -> 0.5105 ms;
				 * round 2 sweep, profiles/r02a_variant_sweep.txt: 0.5040 -> 0.4991 ms); on since round 2 (OCT_R2_NOSHIFT) */
				if ((R == 1 || OCT_R2_NOSHIFT) && sh == 0) store8(cvt0); else store8(cvt);
				/* the remaining HB + HA halo samples (Lanczos only) */
				for (int q4 = N / 4 + tig; q4 < SE / 4; q4 += 32 * R) f4[q4] = cvt(s2[q4]);
				if constexpr (SA == SA_CUBIC && !ROLL) {
					/* mirrored first tap of the cubic: f[-1] = f[1] (cuda_code.cu:284) */
					if (tig == 0) fslot[SPLIT ? SPLIT_ODD_BASE - 1 : HBv - 1] = (float)(reinterpret_cast<const uint16_t*>(slot)[HBv + 1] >> sh);
				}
			}
			if constexpr (ROLL) {
				/* ---- rolling-mean background (cuda_code.cu:165-211): exact integer prefix sums, windows clipped per line ---- */
				if (p == 0) {
					const uint16_t* s16 = reinterpret_cast<const uint16_t*>(slot);
					unsigned carry = 0;
					if (lane == 0) prefix[0] = 0;
					for (int c = 0; c < SE; c += 32) {
						const int q = c + lane;
						unsigned x = (q < SE) ? ((unsigned)s16[q] >> a.shiftBits) : 0u;
#pragma unroll
						for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
						if (q < SE) prefix[q + 1] = carry + x;
						carry += __shfl_sync(0xffffffffu, x, 31);
					}
				}
			}
			group_sync<R>(barId);
			/* raw slot consumed: refill it with the next line of this group */
			if (tig == 0 && glineNext < a.lines) issue_line_load<R, SA == SA_LANCZOS, SRC>(a, glineNext, slot, bar);
			if constexpr (ROLL) {
				const int W = a.W;
				for (int q = tig; q < SE; q += 32 * R) {
					int lo, hi;
					if (q < HBv) { lo = 0; hi = HBv - 1; }
					else if (q >= HBv + N) { lo = HBv + N; hi = SE - 1; }
					else { lo = HBv; hi = HBv + N - 1; }
					const int s = max(lo, q - W + 1), e = min(hi, q + W);
					const float mean = __fdividef((float)(prefix[e + 1] - prefix[s]), (float)(e - s + 1));
					fslot[q] -= mean;
				}
				group_sync<R>(barId);
				if constexpr (SA == SA_CUBIC) {
					if (tig == 0) fslot[HBv - 1] = fslot[HBv + 1];
					group_sync<R>(barId);
				}
			}

			const float* f = fslot + HBv;
			/* the reference clamps the Lanczos line offset to >= 8 (cuda_code.cu:313): line 0 of the buffer is read 8 samples late */
			const int shift = (SA == SA_LANCZOS && gline == 0) ? 8 : 0;
#if OCT_TMEM_LUT
			stage_a_tmem<SA, R>(lane, p, f, shift, tq, v);
#else
			stage_a<SA, R>(lane, p, f, shift, sB, v);
#endif
			group_sync<R>(barId);          /* all gathers done before the exchange tile (aliasing the slot) is written */
		} else {
			const float2* in = a.cin + (size_t)gline * N;
#pragma unroll
			for (int j = 0; j < 32; ++j) v[j] = __ldg(in + R * (lane + 32 * j) + p);
		}

		/* ---- 1024-point inverse FFT of this warp's sub-sequence (first butterfly stage already done by the 4-tap stage A) ---- */
		fft32_inv<src_is_raw(SRC) && stage_a_fuses_stage0(SA)>(v);
#if OCT_TMEM_LUT
		exchange_store_tmem<R>(lane, v, tile, tq);
#else
		exchange_store(lane, v, tile, sTw);
#endif
		__syncwarp();
		exchange_load(lane, v, tile);
		fft32_inv(v);

		if constexpr (R == 2) {
			__syncwarp();                  /* own tile fully read before it is reused for the hand-over */
#if OCT_TMEM_LUT
			combine_store_tmem(lane, p, v, tile, tq);
#else
			combine_store(lane, p, v, tile, sCtw);
#endif
			group_sync<R>(barId);
			combine_load(lane, p, v, partnerTile);
		}

		/* ---- epilogue ---- */
		if (a.cplxOut != nullptr) {
			float2* o = a.cplxOut + (size_t)gline * H;
			if (R == 1 || p == 0) epilogue_complex<0>(lane, v, o); else epilogue_complex<16>(lane, v, o);
		} else {
			int b = gline / a.A, al = gline - b * a.A;
			if (a.flip && (((unsigned)b + a.bscanBase) & 1u) == 0u && (unsigned)b + a.bscanBase < a.flipEnd) al = a.A - 1 - al;
			float* o = a.out + (size_t)blockIdx.y * a.trialOutStride + ((size_t)b * a.A + al) * H;
#if OCT_TMEM_LUT
			ConvOut co;
			co.line = CONV ? a.convOut + ((size_t)b * a.A + al) * H : nullptr;
			co.scale = a.convScale;
			if (R == 1 || p == 0) epilogue_tmem<R, 0, CONV>(lane, v, a.epi, tq, o, co); else epilogue_tmem<R, 16, CONV>(lane, v, a.epi, tq, o, co);
#else
			static_assert(!CONV, "the converted output is written by the tensor-memory epilogue only (OCT_TMEM_LUT = 1)");
			if (R == 1 || p == 0) epilogue_scaled<0>(lane, v, a.epi, sMean, sPpbg, o); else epilogue_scaled<16>(lane, v, a.epi, sMean, sPpbg, o);
#endif
		}
		group_sync<R>(barId);              /* tile / slot reads finished before the next line's conversion overwrites them */
#if OCT_TMEM_LUT
		/* ---- en-face gather: once a block of LB consecutive lines is complete, lanes 0 .. cnt-1 of the group's first warp read the
		 * displayed depth bin of "their" line back (the stores above are ordered before the barrier; the lines are still in L2) and
		 * store it into every rank's frame -- neighbouring lines are neighbouring frame elements, so the peer stores coalesce.  The
		 * epilogue itself carries no per-output cost for the gather. ---- */
		if (egK2 >= 0 && a.cplxOut == nullptr && (R == 1 || p == 0) && (jb + 1 == LB || glineNext >= a.lines)) {
			if (lane <= jb) {
				const int l = gline - jb + lane;
				int b = l / a.A, al = l - b * a.A;
				if (a.flip && (((unsigned)b + a.bscanBase) & 1u) == 0u && (unsigned)b + a.bscanBase < a.flipEnd) al = a.A - 1 - al;
				const size_t row = (size_t)b * a.A + al;
				const float val = __ldcg(a.out + row * H + a.eg.frameNr);
				const unsigned idx = (a.eg.Eglobal - 1u) - (a.eg.offset + (unsigned)row);      /* the reference writes the frame reversed (cuda_code.cu:909) */
#pragma unroll 1
				for (int r = 0; r < a.eg.world; ++r) a.eg.frames[r][idx] = val;
			}
		}
#endif
		gline = glineNext;
		jb = (jb + 1 == LB) ? 0 : jb + 1;
	}
#if OCT_TMEM_LUT
	tmem_fence_before_sync();
	__syncthreads();
	if (warp == 0) tmem_dealloc(tmemBase, TmemMap<R>::ALLOC);
#else
	__syncthreads();
#endif
	if (a.eg.world > 0) gather_publish(a.eg);      /* behind the CTA-wide barrier above */

}

}  // namespace octb200
