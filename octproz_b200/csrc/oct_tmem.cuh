/*
 * oct_tmem.cuh -- tensor memory (TMEM, 256 KB per SM) as a per-lane constant store for the fused kernel.
 *
 * The fused kernel is bounded by shared-memory bandwidth (128 B/clk/SM), and a third of that traffic is tables that
 * are identical for every A-scan and private to a lane: lane l always handles samples l + 32 j, bins l + 32 k2 and the
 * twiddles w^{l k}.  TMEM is organised as 128 lanes x 512 32-bit columns, a thread reads N consecutive columns of ITS lane
 * with tcgen05.ld.32x32b.xN over a datapath that is separate from shared memory (measured here with tools/tmem_bench.cu:
 * > 360 B/clk/SM, 3x shared memory).  So the tables live in TMEM: written once per launch with tcgen05.st (the same
 * image into each of the four lane quadrants, warp w reads quadrant w % 4), read per line with tcgen05.ld.
 * No tensor-core MMA is involved -- TMEM is used as what it physically is, a big lane-private register-file extension.
 *
 * Column map of one lane quadrant (512 columns allocated):
 *   [0,256)   stage LUT: 16 row pairs (rows j and j + 16) x 16 words { off_a off_b w0a w0b w1a w1b t_a t_b | w2a w2b w3a w3b wP_a wP_b }
 *   [256,320) inter-pass twiddles w^{k1 lane}, k1 = 0..31 (32 complex)
 *   [320,384) R = 2 only: combine twiddles w_2048^{lane + 32 k2}, k2 = 0..31
 *   MEAN      FPN mean line, 16 complex of the bins this warp finalises     PPBG  background, 16 floats
 * For R = 2 (two warps per line, p = warp % 2) a quadrant holds only the tables of its own p (quadrant = warp % 4 -> p = quadrant & 1).
 */
#pragma once
#include "oct_device.cuh"

namespace octb200 {

#ifndef OCT_R2_EGVAR
#define OCT_R2_EGVAR 0
#endif
/* the final scale FMA of two outputs as one packed instruction (same FMA per half, bit-identical): 0.2042 -> 0.2038 ms (N = 1024),
 * 0.5030 -> 0.4997 ms (N = 2048), profiles/r02b_variant_sweep.txt */
#ifndef OCT_EPI_PAIR
#define OCT_EPI_PAIR 1
#endif
/* Inter-pass twiddle layout and read width, per line-group size R (measured on one B200 box, tools/variant_sweep.sh,
 * profiles/r02a_variant_sweep.txt; every variant is bit-identical, same output hash):
 *   TW4  : every inter-pass twiddle is stored as four words (t.x, t.y, -t.y, t.x), the two operand pairs of the packed complex multiply,
 *          instead of two -- a half-pair negation is not a free operand modifier, so the two-word form costs one FADD per twiddle and
 *          line (31).  Same arithmetic bit for bit.
 *   XCHG : registers per tensor-memory read of the exchange phase.  8 = double-buffered x8 reads (ptxas hoists them and copies the
 *          still-live tuple out of the way: 66 MOV per line); 32 = two single-buffered x32 reads whose destination registers are the
 *          ones the stored values free up (MOV 74 -> 23 per line).
 *   N = 1024 (R = 1): TW4 + x32 0.2114 -> 0.2042 ms per 1024x512x256 volume (tw4 alone 0.2072, x32 alone 0.2094).
 *   N = 2048 (R = 2): x32 costs the two-warp kernel 1.5-4 % (0.504 -> 0.512-0.524 ms per 2048x1024x128 buffer); tw4 alone 0.4998,
 *                     tw4 + x16 0.4994 (profiles/r02b_variant_sweep.txt; without tw4 0.5127). */
#ifndef OCT_R1_TW4
#define OCT_R1_TW4 1
#endif
#ifndef OCT_R2_TW4
#define OCT_R2_TW4 1
#endif
#ifndef OCT_R1_XCHG_X
#define OCT_R1_XCHG_X 32
#endif
#ifndef OCT_R2_XCHG_X
#define OCT_R2_XCHG_X 16
#endif
template <int R> struct XchgCfg {
	static constexpr bool TW4 = (R == 1) ? (OCT_R1_TW4 != 0) : (OCT_R2_TW4 != 0);
	static constexpr int X = (R == 1) ? OCT_R1_XCHG_X : OCT_R2_XCHG_X;
};
template <int R> struct TmemMap;
/* per lane quadrant; for R = 2 a quadrant only holds the tables of ITS sub-sequence p = quadrant & 1 (warp w: p = w % 2, quadrant w % 4) */
template <> struct TmemMap<1> {
	static constexpr int LUT = 0, TW = 256, CTW = XchgCfg<1>::TW4 ? 384 : 320, MEAN = CTW, PPBG = MEAN + 32, ALLOC = 512;
};
template <> struct TmemMap<2> {
	static constexpr int LUT = 0, TW = 256, CTW = XchgCfg<2>::TW4 ? 384 : 320, MEAN = CTW + 64, PPBG = MEAN + 32, ALLOC = 512;
};

/* ---- raw tcgen05 wrappers (SASS: LDTM / STTM / UTCALLOC) ---- */
__device__ __forceinline__ void tmem_alloc(uint32_t* smemResult, int cols) {
	if (cols == 256) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(smemResult)) : "memory");
	else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smemResult)) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, int cols) {
	if (cols == 256) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(taddr) : "memory");
	else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
	             ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

/* loads: the wait carries the destination registers as read-write operands so no consumer can be scheduled before it */
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;"
	             : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	             : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
	               "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;"
	             : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
	               "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]) :: "memory");
}
/* split issue / wait: the read of the NEXT table row is in flight while the current row is being used.  The wait names the
 * destination registers as read-write operands, so every consumer is ordered behind it; nothing may touch them in between. */
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&r)[16]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	             : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
	               "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(float (&r)[16]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;"
	             : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
	               "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]) :: "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, float (&r)[8]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_wait(float (&r)[8]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;"
	             : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld8_wait2(float (&r)[8], float (&s)[8]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;"
	             : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),
	               "+f"(s[0]), "+f"(s[1]), "+f"(s[2]), "+f"(s[3]), "+f"(s[4]), "+f"(s[5]), "+f"(s[6]), "+f"(s[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, float (&r)[4]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4_wait(float (&r)[4]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_ld2_issue(uint32_t taddr, float (&r)[2]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld2_wait(float (&r)[2]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]) :: "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&r)[32]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
	             : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]),
	               "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]),
	               "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(float (&r)[32]) {
	asm volatile("tcgen05.wait::ld.sync.aligned;"
	             : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]),
	               "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]),
	               "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]), "+f"(r[30]), "+f"(r[31]) :: "memory");
}
/* width picked by the array type */
__device__ __forceinline__ void tmem_ldx_issue(uint32_t taddr, float (&r)[32]) { tmem_ld32_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldx_issue(uint32_t taddr, float (&r)[16]) { tmem_ld16_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldx_wait(float (&r)[32]) { tmem_ld32_wait(r); }
__device__ __forceinline__ void tmem_ldx_wait(float (&r)[16]) { tmem_ld16_wait(r); }
__device__ __forceinline__ void tmem_ldx_issue(uint32_t taddr, float (&r)[8]) { tmem_ld8_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldx_issue(uint32_t taddr, float (&r)[4]) { tmem_ld4_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldx_issue(uint32_t taddr, float (&r)[2]) { tmem_ld2_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldx_wait(float (&r)[8]) { tmem_ld8_wait(r); }
__device__ __forceinline__ void tmem_ldx_wait(float (&r)[4]) { tmem_ld4_wait(r); }
__device__ __forceinline__ void tmem_ldx_wait(float (&r)[2]) { tmem_ld2_wait(r); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&r)[4]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&r)[2]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]) :: "memory");
}

/* ---- fill: warp q (< 4) writes the image of lane `lane` into quadrant q ---- */
__device__ __forceinline__ void tmem_st_f4(uint32_t taddr, float4 a) {
	const uint32_t r[4] = { __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w) };
	tmem_st4(taddr, r);
}
/* Every warp takes part: warp w writes rows part, part + parts, ... of quadrant w % 4 (part = w / 4, parts = warps of that quadrant).
 * All global loads of a phase are issued before the first tensor-memory store, so the fill costs two memory latencies instead of
 * one per row (it used to be ~7 % of the kernel: 50 dependent L2 round trips by 4 warps while 12 waited at the barrier). */
template <int R>
__device__ __forceinline__ void tmem_fill(uint32_t tq /* quadrant base */, int quadrant, int part, int parts, int lane, const FusedArgs& a, const float4* lutB, bool haveLut) {
	using M = TmemMap<R>;
	constexpr int HN = 512 * R;
	const int p = (R == 2) ? (quadrant & 1) : 0;         /* the sub-sequence whose warps read this quadrant */
	if (haveLut) {
		for (int base = part; base < 16; base += 4 * parts) {
			float4 b[4][4];
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int jj = base + u * parts;
				if (jj < 16) {
					const int e = p * 512 + lane + 32 * jj;
					b[u][0] = __ldg(lutB + e); b[u][1] = __ldg(lutB + HN + e); b[u][2] = __ldg(lutB + 2 * HN + e); b[u][3] = __ldg(lutB + 3 * HN + e);
				}
			}
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const int jj = base + u * parts;
				if (jj < 16) {
					const float4 Pq = b[u][0], Qq = b[u][1], W01 = b[u][2], W23 = b[u][3];
					/* row = [ off_a off_b w0a w0b | w1a w1b t_a t_b || w2a w2b w3a w3b | wPa wPb ]: two x8 reads, the second one late */
					tmem_st_f4(tq + M::LUT + 16 * jj + 0, make_float4(Qq.x, Qq.y, W01.x, W01.y));
					tmem_st_f4(tq + M::LUT + 16 * jj + 4, make_float4(W01.z, W01.w, Qq.z, Qq.w));
					tmem_st_f4(tq + M::LUT + 16 * jj + 8, W23);
					tmem_st_f4(tq + M::LUT + 16 * jj + 12, Pq);
				}
			}
		}
	}
	const int k2lo = 16 * p;                              /* bins lane + 32 k2, k2 in [16p, 16p+16) are finalised by the warps of p */
	const bool fpn = a.epi.fpn && a.cplxOut == nullptr;
	for (int base = part; base < 16; base += 4 * parts) {
		float4 t[4], c[4], m[4], g[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const int i = base + u * parts;                 /* pair index: twiddles k1 = 2i, 2i+1; mean bins 2i, 2i+1 (i < 8); background bins 4i..4i+3 (i < 4) */
			if (i < 16) {
				const float2 t0 = __ldg(a.tw + (2 * i) * 32 + lane), t1 = __ldg(a.tw + (2 * i + 1) * 32 + lane);
				t[u] = make_float4(t0.x, t0.y, t1.x, t1.y);
				if constexpr (R == 2) {
					const float2 c0 = __ldg(a.ctw + lane + 32 * (2 * i)), c1 = __ldg(a.ctw + lane + 32 * (2 * i + 1));
					c[u] = make_float4(c0.x, c0.y, c1.x, c1.y);
				}
				if (fpn && i < 8) {
					const float2 m0 = __ldg(a.meanLine + lane + 32 * (k2lo + 2 * i)), m1 = __ldg(a.meanLine + lane + 32 * (k2lo + 2 * i + 1));
					m[u] = make_float4(m0.x, m0.y, m1.x, m1.y);
				}
				if (a.epi.ppbg && i < 4)
					g[u] = make_float4(__ldg(a.ppbg + lane + 32 * (k2lo + 4 * i)), __ldg(a.ppbg + lane + 32 * (k2lo + 4 * i + 1)),
					                   __ldg(a.ppbg + lane + 32 * (k2lo + 4 * i + 2)), __ldg(a.ppbg + lane + 32 * (k2lo + 4 * i + 3)));
			}
		}
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const int i = base + u * parts;
			if (i < 16) {
				if constexpr (XchgCfg<R>::TW4) {
					tmem_st_f4(tq + M::TW + 8 * i, make_float4(t[u].x, t[u].y, -t[u].y, t[u].x));
					tmem_st_f4(tq + M::TW + 8 * i + 4, make_float4(t[u].z, t[u].w, -t[u].w, t[u].z));
				} else {
					tmem_st_f4(tq + M::TW + 4 * i, t[u]);
				}
				if constexpr (R == 2) tmem_st_f4(tq + M::CTW + 4 * i, c[u]);
				if (fpn && i < 8) tmem_st_f4(tq + M::MEAN + 4 * i, m[u]);
				if (a.epi.ppbg && i < 4) tmem_st_f4(tq + M::PPBG + 4 * i, g[u]);
			}
		}
	}
	tmem_wait_st();
}

/* ---- stage A from TMEM (same arithmetic as stage_a in oct_phases.cuh): two x8 reads per row pair ---- */
template <int SA, int R>
__device__ __forceinline__ void stage_a_tmem(int lane, int p, const float* f, int shift, uint32_t tq, float2 (&v)[32]) {
	using M = TmemMap<R>;
	if constexpr (SA == SA_CUBIC || SA == SA_LINEAR) {
		/* 4-tap interpolators.  Table row pair = [ offX_a offX_b w0a w0b | w1a w1b offY_a offY_b || w2a w2b w3a w3b | wPa wPb ]:
		 * the taps of a sample are f[offX], f[offX+4], f[offY], f[offY+4] with weights w0..w3 (natural slot: offY = offX + 8;
		 * parity-split slot of R = 2: one pair from the even half, one from the odd half, weights permuted on the host).
		 * Two x8 reads per row pair at 128 registers per thread: the first half of the NEXT row pair (tap offsets, first weights)
		 * is prefetched while the current one is gathered; the second half (late weights, window x phasor) is read while the
		 * gathers are in flight.  (One x16 read per row pair with the next one prefetched needs more contiguous registers than
		 * ptxas finds: it spills and was measured slower.) */
		float qa[2][8];
		tmem_ld8_issue(tq + M::LUT, qa[0]);
		tmem_ld8_wait(qa[0]);
		static_for<0, 16>([&](auto jc) {
			constexpr int jj = decltype(jc)::value;
			constexpr int c = jj & 1;
			const int xa = __float_as_int(qa[c][0]), xb = __float_as_int(qa[c][1]), ya = __float_as_int(qa[c][6]), yb = __float_as_int(qa[c][7]);
			const float2 Y0 = make_float2(ldf(f, xa), ldf(f, xb)), Y1 = make_float2(ldf(f, xa + 4), ldf(f, xb + 4));
			const float2 Y2 = make_float2(ldf(f, ya), ldf(f, yb)), Y3 = make_float2(ldf(f, ya + 4), ldf(f, yb + 4));
			float w[8];
			tmem_ld8_issue(tq + M::LUT + 16 * jj + 8, w);
			if constexpr (jj < 15) tmem_ld8_issue(tq + M::LUT + 16 * (jj + 1), qa[c ^ 1]);
			float2 y = pfma(make_float2(qa[c][4], qa[c][5]), Y1, pmul(make_float2(qa[c][2], qa[c][3]), Y0));
			if constexpr (jj < 15) tmem_ld8_wait2(w, qa[c ^ 1]); else tmem_ld8_wait(w);
			y = pfma(make_float2(w[2], w[3]), Y3, pfma(make_float2(w[0], w[1]), Y2, y));
			const float2 t = cscale(make_float2(w[4], w[5]), y.x);          /* first FFT butterfly folded in, see sample_taps4_x2 */
			v[jj] = pfma(make_float2(w[6], w[7]), make_float2(y.y, y.y), t);
			v[jj + 16] = pfma(make_float2(w[6], w[7]), make_float2(-y.y, -y.y), t);
		});
		return;
	}
#pragma unroll
	for (int jj = 0; jj < 16; ++jj) {
		float q[8], w[8];
		tmem_ld8(tq + M::LUT + 16 * jj, q);                 /* off_a off_b . . . . t_a t_b */
		tmem_ld8(tq + M::LUT + 16 * jj + 8, w);             /* . . . . wPa.x wPa.y wPb.x wPb.y */
		const float2 wa = make_float2(w[4], w[5]), wb = make_float2(w[6], w[7]);
		if constexpr (SA == SA_NONE) {
			const int s = lane + 32 * jj;
			v[jj] = cscale(wa, f[R * s + p]);
			v[jj + 16] = cscale(wb, f[R * (s + 512) + p]);
		} else {
			v[jj] = sample_lanczos(f, shift, make_float4(q[0], wa.x, wa.y, q[6]));
			v[jj + 16] = sample_lanczos(f, shift, make_float4(q[1], wb.x, wb.y, q[7]));
		}
	}
}

/* ---- inter-pass twiddle + transpose store, twiddles from TMEM (cf. exchange_store): w^{k1 lane} at columns TW + 2 k1 ----
 * XchgCfg<R>::X = registers per tensor-memory read (see the top of this file).  With X = 8 the read of chunk c+1 is in flight while
 * chunk c is multiplied and stored. */
template <int R>
__device__ __forceinline__ void exchange_store_tmem(int lane, const float2 (&v)[32], float2* xbuf, uint32_t tq) {
	using M = TmemMap<R>;
	constexpr bool TW4 = XchgCfg<R>::TW4;
	constexpr int WPT = TW4 ? 4 : 2;                              /* words per twiddle */
	constexpr int X = XchgCfg<R>::X, TW = X / WPT, CH = 32 / TW;  /* registers per read, twiddles per read, reads per line */
	auto twmul = [](float2 val, const float* w) {
		if constexpr (TW4) return pfma(make_float2(val.y, val.y), make_float2(w[2], w[3]), pmul(make_float2(val.x, val.x), make_float2(w[0], w[1])));
		else return cmul(val, make_float2(w[0], w[1]));
	};
	auto issue = [&](int c, float (&dst)[X]) { tmem_ldx_issue(tq + M::TW + X * c, dst); };
	auto wait = [&](float (&dst)[X]) { tmem_ldx_wait(dst); };
	if constexpr (X >= 16) {
		/* wide reads, not double buffered: the twiddles of a whole half (or all) of the exchange at once, their registers are the ones
		 * the stored values free up */
		static_for<0, CH>([&](auto cc) {
			constexpr int c = decltype(cc)::value;
			float t1[X];
			issue(c, t1);
			wait(t1);
			static_for<0, TW>([&](auto ic) {
				constexpr int i = decltype(ic)::value;
				constexpr int k1 = TW * c + i;
				constexpr int r = bitrev5(k1);
				float2 val = v[r];
				if constexpr (k1 != 0) val = twmul(val, &t1[WPT * i]);
				xbuf[k1 * XPITCH + lane] = val;
			});
		});
		return;
	}
	float t[2][X];
	issue(0, t[0]);
	wait(t[0]);
	static_for<0, CH>([&](auto cc) {
		constexpr int c = decltype(cc)::value;            /* k1 = TW c .. TW c + TW - 1 */
		constexpr int b = c & 1;
		if constexpr (c < CH - 1) issue(c + 1, t[b ^ 1]);
		static_for<0, TW>([&](auto ic) {
			constexpr int i = decltype(ic)::value;
			constexpr int k1 = TW * c + i;
			constexpr int r = bitrev5(k1);
			float2 val = v[r];
			if constexpr (k1 != 0) val = twmul(val, &t[b][WPT * i]);
			xbuf[k1 * XPITCH + lane] = val;
		});
		if constexpr (c < CH - 1) wait(t[b ^ 1]);
	});
}

/* ---- R = 2 combine with the combine twiddles from TMEM (cf. combine_store) ---- */
__device__ __forceinline__ void combine_store_tmem(int lane, int p, float2 (&v)[32], float2* ownTile, uint32_t tq) {
	using M = TmemMap<2>;
	static_for<0, 16>([&](auto cc) {
		constexpr int c = decltype(cc)::value;           /* k2 = 2c, 2c+1 */
		float t[4];
		if (p == 1) tmem_ld4(tq + M::CTW + 4 * c, t);
		static_for<0, 2>([&](auto ec) {
			constexpr int k2 = 2 * c + decltype(ec)::value;
			constexpr int r = bitrev5(k2);
			if (p == 1) v[r] = cmul(v[r], make_float2(t[2 * decltype(ec)::value], t[2 * decltype(ec)::value + 1]));
			if constexpr (k2 < 16) { if (p == 1) ownTile[k2 * 32 + lane] = v[r]; }
			else                   { if (p == 0) ownTile[(k2 - 16) * 32 + lane] = v[r]; }
		});
	});
}

/* ---- epilogue with the FPN line / background from TMEM (cf. epilogue_scaled_t) ----
 * CONV: floatToOutput (cuda_code.cu:943-967) into u16 containers in the same pass (the fused kernel only takes u16 raw data, and the
 * output container is the input's).  The reference computes
 * (container)((double)saturate(x) * K), K = 2^bits - 1, i.e. floor of the exact product.  One round-toward-zero FMA gives the same
 * integer without the conversion pipe: saturate(x) * K + 2^23 rounded toward zero is exactly 2^23 + floor(saturate(x) * K)
 * (the product is >= 0 and K < 2^23, so the ulp of the sum is 1), and its low mantissa bits are the container value. */
struct ConvOut {
	unsigned short* line;    /* converted line (same bin order as outLine); unused without CONV */
	float scale;
};
template <int R, int K2LO, bool LOG, bool FPN, bool PPBG, bool CONV, bool EG>
__device__ __forceinline__ void epilogue_tmem_t(int lane, const float2 (&v)[32], const EpiConsts& e, uint32_t tq, float* outLine, const ConvOut& co, int egK2, float& egVal) {
	using M = TmemMap<R>;
	const float sA = e.scaleA, sB = e.scaleB, bw = e.ppbgWeight, bo = e.ppbgOffset;
	static_for<0, 4>([&](auto gc) {
		constexpr int g = decltype(gc)::value;           /* four bins per TMEM read: k2 = K2LO + 4g .. +3 */
		float m[8], bgv[4];
		if constexpr (FPN) tmem_ld8(tq + M::MEAN + 8 * g, m);
		if constexpr (PPBG) tmem_ld4(tq + M::PPBG + 4 * g, bgv);
#if OCT_EPI_PAIR
		/* experiment: the final scale FMA of two outputs as one packed instruction (same FMA per half, bit-identical) */
		static_for<0, 2>([&](auto hc) {
			constexpr int h = decltype(hc)::value;
			float tr[2];
			static_for<0, 2>([&](auto ic) {
				constexpr int i = 2 * h + decltype(ic)::value;
				constexpr int r = bitrev5(K2LO + 4 * g + i);
				float2 d = v[r];
				if constexpr (FPN) d = csub(d, make_float2(m[2 * i], m[2 * i + 1]));
				const float pw = fmaf(d.x, d.x, d.y * d.y);
				tr[decltype(ic)::value] = LOG ? oct_lg2(pw) : oct_sqrt(pw);
			});
			const float2 o2 = pfma(make_float2(tr[0], tr[1]), make_float2(sA, sA), make_float2(sB, sB));
			static_for<0, 2>([&](auto ic) {
				constexpr int i = 2 * h + decltype(ic)::value;
				constexpr int k2 = K2LO + 4 * g + i;
				const int z = lane + 32 * k2;
				float o = decltype(ic)::value ? o2.y : o2.x;
				if constexpr (PPBG) o = saturate01(o - fmaf(bw, bgv[i], bo));
				outLine[z] = o;
				if constexpr (CONV) co.line[z] = (unsigned short)__float_as_uint(__fmaf_rz(__saturatef(o), co.scale, 8388608.0f));
				if constexpr (EG) { if (k2 == egK2) egVal = o; }
			});
		});
#else
		static_for<0, 4>([&](auto ic) {
			constexpr int i = decltype(ic)::value;
			constexpr int k2 = K2LO + 4 * g + i;
			constexpr int r = bitrev5(k2);
			const int z = lane + 32 * k2;
			float2 d = v[r];
			if constexpr (FPN) d = csub(d, make_float2(m[2 * i], m[2 * i + 1]));
			const float pw = fmaf(d.x, d.x, d.y * d.y);
			float o = LOG ? fmaf(oct_lg2(pw), sA, sB) : fmaf(oct_sqrt(pw), sA, sB);
			if constexpr (PPBG) o = saturate01(o - fmaf(bw, bgv[i], bo));
			outLine[z] = o;
			if constexpr (CONV) co.line[z] = (unsigned short)__float_as_uint(__fmaf_rz(__saturatef(o), co.scale, 8388608.0f));
			if constexpr (EG) { if (k2 == egK2) egVal = o; }   /* uniform compare: the displayed en-face bin stays in a register */
		});
#endif
	});
}
/* runtime flags -> one uniform branch per line instead of several per output.  EG (the en-face gather keeps one output per line in a
 * register) is a variant of its own: without it the per-output compare + select would cost 32 instructions per line for nothing */
template <int R, int K2LO, bool CONV, bool EG>
__device__ __forceinline__ void epilogue_tmem_sel(int lane, const float2 (&v)[32], const EpiConsts& e, uint32_t tq, float* outLine, const ConvOut& co, int egK2, float& egVal) {
	const int sel = (e.logMode ? 1 : 0) | (e.fpn ? 2 : 0) | (e.ppbg ? 4 : 0);
	switch (sel) {
	case 0: epilogue_tmem_t<R, K2LO, false, false, false, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	case 1: epilogue_tmem_t<R, K2LO, true, false, false, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	case 2: epilogue_tmem_t<R, K2LO, false, true, false, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	case 3: epilogue_tmem_t<R, K2LO, true, true, false, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	case 4: epilogue_tmem_t<R, K2LO, false, false, true, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	case 5: epilogue_tmem_t<R, K2LO, true, false, true, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	case 6: epilogue_tmem_t<R, K2LO, false, true, true, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	default: epilogue_tmem_t<R, K2LO, true, true, true, CONV, EG>(lane, v, e, tq, outLine, co, egK2, egVal); break;
	}
}
/* (EG: the variant that keeps the displayed en-face bin in a register while storing is no longer used by the fused kernel -- the gather
 * reads the finished values back, one block of lines at a time, k_fused.cuh -- so the epilogue has no per-output cost for it) */
template <int R, int K2LO, bool CONV>
__device__ __forceinline__ void epilogue_tmem(int lane, const float2 (&v)[32], const EpiConsts& e, uint32_t tq, float* outLine, const ConvOut& co) {
	float unused = 0.f;
	epilogue_tmem_sel<R, K2LO, CONV, false>(lane, v, e, tq, outLine, co, -1, unused);
}

}  // namespace octb200
