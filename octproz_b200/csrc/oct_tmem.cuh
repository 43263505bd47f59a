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
 * Column map (R = N/1024 warps per line):
 *   R = 1: [0,128) stage LUT (16 row pairs x {P quad, Q quad})  [128,148) twiddles [A1 B1][A2 A3][A4 A5][A6 A7][B2 B3]  [148,180) FPN mean (16 cplx)
 *          [180,196) background (16 floats)                                               -> 256 columns allocated
 *   R = 2: [0,256) stage LUT of p=0,1  [256,276) twiddles  [276,340) combine twiddles (32 cplx)  [340,404) mean (2 x 16 cplx)
 *          [404,436) background (2 x 16)                                                   -> 512 columns allocated
 */
#pragma once
#include "oct_device.cuh"

namespace octb200 {

template <int R> struct TmemMap;
template <> struct TmemMap<1> { static constexpr int LUT = 0, TW = 128, CTW = 148, MEAN = 148, PPBG = 180, ALLOC = 256; };
template <> struct TmemMap<2> { static constexpr int LUT = 0, TW = 256, CTW = 276, MEAN = 340, PPBG = 404, ALLOC = 512; };

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
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&r)[4]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]) :: "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&r)[2]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(taddr));
	asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]) :: "memory");
}

/* ---- fill: warp q (< 4) writes the image of lane `lane` into quadrant q ---- */
template <int R>
__device__ __forceinline__ void tmem_fill(uint32_t tq /* quadrant base */, int lane, const FusedArgs& a, bool haveLut) {
	using M = TmemMap<R>;
	constexpr int N = 1024 * R, H = N / 2;
	if (haveLut) {
		const float4* P = a.lutB;
		const float4* Q = a.lutB + H;
		for (int p = 0; p < R; ++p)
			for (int jj = 0; jj < 16; ++jj) {
				const float4 Pq = __ldg(P + p * 512 + lane + 32 * jj), Qq = __ldg(Q + p * 512 + lane + 32 * jj);
				const uint32_t r[8] = { __float_as_uint(Pq.x), __float_as_uint(Pq.y), __float_as_uint(Pq.z), __float_as_uint(Pq.w),
				                        __float_as_uint(Qq.x), __float_as_uint(Qq.y), __float_as_uint(Qq.z), __float_as_uint(Qq.w) };
				tmem_st8(tq + M::LUT + p * 128 + 8 * jj, r);
			}
	}
	{   /* twiddles (oct_luts.hpp build_twiddles_1024) in the order [A1 B1][A2 A3][A4 A5][A6 A7][B2 B3]: 20 columns, 5 x4 stores */
		float2 t[10];
		t[0] = __ldg(a.tw + 1 * 32 + lane); t[1] = __ldg(a.tw + 256 + 1 * 32 + lane);
		for (int i = 2; i < 8; ++i) t[i] = __ldg(a.tw + i * 32 + lane);
		t[8] = __ldg(a.tw + 256 + 2 * 32 + lane); t[9] = __ldg(a.tw + 256 + 3 * 32 + lane);
		for (int c = 0; c < 5; ++c) {
			const uint32_t r[4] = { __float_as_uint(t[2 * c].x), __float_as_uint(t[2 * c].y), __float_as_uint(t[2 * c + 1].x), __float_as_uint(t[2 * c + 1].y) };
			tmem_st4(tq + M::TW + 4 * c, r);
		}
	}
	if constexpr (R == 2) {
		for (int k2 = 0; k2 < 32; k2 += 2) {
			const float2 c0 = __ldg(a.ctw + lane + 32 * k2), c1 = __ldg(a.ctw + lane + 32 * (k2 + 1));
			const uint32_t r[4] = { __float_as_uint(c0.x), __float_as_uint(c0.y), __float_as_uint(c1.x), __float_as_uint(c1.y) };
			tmem_st4(tq + M::CTW + 2 * k2, r);
		}
	}
	if (a.epi.fpn && a.cplxOut == nullptr) {
		for (int k2 = 0; k2 < 16 * R; k2 += 2) {
			const float2 m0 = __ldg(a.meanLine + lane + 32 * k2), m1 = __ldg(a.meanLine + lane + 32 * (k2 + 1));
			const uint32_t r[4] = { __float_as_uint(m0.x), __float_as_uint(m0.y), __float_as_uint(m1.x), __float_as_uint(m1.y) };
			tmem_st4(tq + M::MEAN + 2 * k2, r);
		}
	}
	if (a.epi.ppbg) {
		for (int k2 = 0; k2 < 16 * R; k2 += 4) {
			uint32_t r[4];
			for (int i = 0; i < 4; ++i) r[i] = __float_as_uint(__ldg(a.ppbg + lane + 32 * (k2 + i)));
			tmem_st4(tq + M::PPBG + k2, r);
		}
	}
	tmem_wait_st();
}

/* ---- stage A from TMEM (same arithmetic as stage_a in oct_phases.cuh) ---- */
template <int SA, int R>
__device__ __forceinline__ void stage_a_tmem(int lane, int p, const float* f, int shift, uint32_t tq, float2 (&v)[32]) {
	using M = TmemMap<R>;
	const uint32_t base = tq + M::LUT + p * 128;
#pragma unroll
	for (int jj = 0; jj < 16; ++jj) {
		float q[8];
		tmem_ld8(base + 8 * jj, q);
		const float2 wa = make_float2(q[0], q[1]), wb = make_float2(q[2], q[3]);
		if constexpr (SA == SA_CUBIC) {
			sample_cubic_x2(f, __float_as_int(q[4]), __float_as_int(q[5]), make_float2(q[6], q[7]), wa, wb, v[2 * jj], v[2 * jj + 1]);
		} else if constexpr (SA == SA_LINEAR) {
			sample_linear_x2(f, __float_as_int(q[4]), __float_as_int(q[5]), make_float2(q[6], q[7]), wa, wb, v[2 * jj], v[2 * jj + 1]);
		} else if constexpr (SA == SA_NONE) {
			const int s = lane + 64 * jj;
			v[2 * jj] = cscale(wa, f[R * s + p]);
			v[2 * jj + 1] = cscale(wb, f[R * (s + 32) + p]);
		} else {
			v[2 * jj] = sample_lanczos(f, shift, make_float4(q[4], q[0], q[1], q[6]));
			v[2 * jj + 1] = sample_lanczos(f, shift, make_float4(q[5], q[2], q[3], q[7]));
		}
	}
}

/* ---- inter-pass twiddle + transpose store, twiddles from TMEM (cf. exchange_store) ----
 * TMEM twiddle columns (pairs of complex per x4 read): [A1 B1] [A2 A3] [A4 A5] [A6 A7] [B2 B3] */
template <int R>
__device__ __forceinline__ void exchange_store_tmem(int lane, const float2 (&v)[32], float2* xbuf, uint32_t tq) {
	using M = TmemMap<R>;
	float cb[4];
	tmem_ld4(tq + M::TW + 16, cb);
	const float2 B2 = make_float2(cb[0], cb[1]), B3 = make_float2(cb[2], cb[3]);
	float2 B1 = make_float2(1.0f, 0.0f);
	static_for<0, 4>([&](auto hc) {
		constexpr int h = decltype(hc)::value;          /* a = 2h, 2h+1 */
		float c[4];
		tmem_ld4(tq + M::TW + 4 * h, c);
		float2 Aeven, Aodd;
		if constexpr (h == 0) { Aeven = make_float2(1.0f, 0.0f); Aodd = make_float2(c[0], c[1]); B1 = make_float2(c[2], c[3]); }
		else { Aeven = make_float2(c[0], c[1]); Aodd = make_float2(c[2], c[3]); }
		static_for<0, 2>([&](auto ec) {
			constexpr int e = decltype(ec)::value;
			constexpr int aIdx = 2 * h + e;
			const float2 A = e == 0 ? Aeven : Aodd;
			static_for<0, 4>([&](auto bc) {
				constexpr int b = decltype(bc)::value;
				constexpr int k1 = 4 * aIdx + b;
				constexpr int r = bitrev5(k1);
				float2 val = v[r];
				if constexpr (aIdx == 0 && b == 0) { }
				else if constexpr (aIdx == 0) val = cmul(val, b == 1 ? B1 : (b == 2 ? B2 : B3));
				else if constexpr (b == 0) val = cmul(val, A);
				else val = cmul(val, cmul(A, b == 1 ? B1 : (b == 2 ? B2 : B3)));
				xbuf[k1 * XPITCH + lane] = val;
			});
		});
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

/* ---- epilogue with the FPN line / background from TMEM (cf. epilogue_scaled_t) ---- */
template <int R, int K2LO, bool LOG, bool FPN, bool PPBG>
__device__ __forceinline__ void epilogue_tmem_t(int lane, const float2 (&v)[32], const EpiConsts& e, uint32_t tq, float* outLine) {
	using M = TmemMap<R>;
	const float sA = e.scaleA, sB = e.scaleB, bw = e.ppbgWeight, bo = e.ppbgOffset;
	static_for<0, 4>([&](auto gc) {
		constexpr int g = decltype(gc)::value;           /* four bins per TMEM read: k2 = K2LO + 4g .. +3 */
		float m[8], bgv[4];
		if constexpr (FPN) tmem_ld8(tq + M::MEAN + 2 * (K2LO + 4 * g), m);
		if constexpr (PPBG) tmem_ld4(tq + M::PPBG + (K2LO + 4 * g), bgv);
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
		});
	});
}
template <int R, int K2LO>
__device__ __forceinline__ void epilogue_tmem(int lane, const float2 (&v)[32], const EpiConsts& e, uint32_t tq, float* outLine) {
	const int sel = (e.logMode ? 1 : 0) | (e.fpn ? 2 : 0) | (e.ppbg ? 4 : 0);
	switch (sel) {
	case 0: epilogue_tmem_t<R, K2LO, false, false, false>(lane, v, e, tq, outLine); break;
	case 1: epilogue_tmem_t<R, K2LO, true, false, false>(lane, v, e, tq, outLine); break;
	case 2: epilogue_tmem_t<R, K2LO, false, true, false>(lane, v, e, tq, outLine); break;
	case 3: epilogue_tmem_t<R, K2LO, true, true, false>(lane, v, e, tq, outLine); break;
	case 4: epilogue_tmem_t<R, K2LO, false, false, true>(lane, v, e, tq, outLine); break;
	case 5: epilogue_tmem_t<R, K2LO, true, false, true>(lane, v, e, tq, outLine); break;
	case 6: epilogue_tmem_t<R, K2LO, false, true, true>(lane, v, e, tq, outLine); break;
	default: epilogue_tmem_t<R, K2LO, true, true, true>(lane, v, e, tq, outLine); break;
	}
}

}  // namespace octb200
