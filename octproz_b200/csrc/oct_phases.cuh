/*
 * oct_phases.cuh -- the per-lane phases of the fused OCT kernel, written as
 * __host__ __device__ functions of (lane, register file, shared-memory views) so that the
 * lane/register/bank maps can be executed on the CPU by tests/emu (test-only emulator).
 *
 * What each phase restates of the reference (paths under /root/reference/octproz_project/octproz/src):
 *   stage A   : inputToCufftComplex (cuda_code.cu:109-147, done once per sample in the slot conversion),
 *               klinearization{,Cubic,Lanczos}AndWindowingAndDispersionCompensation (cuda_code.cu:413-489)
 *               and every other branch of the dispatch table (cuda_code.cu:1448-1511) through LUT contents
 *   FFT       : cufftExecC2C CUFFT_INVERSE (cuda_code.cu:1514-1515), as a 32x32 four-step transform
 *   epilogue  : meanALineSubtraction (567-584), postProcessTruncateLog/Lin (699-741), cuda_bscanFlip (787-807,
 *               folded into the store address), postProcessBackgroundRemoval (757-767)
 */
#pragma once
#include "fft_core.cuh"

#include <cmath>
#include <cstring>

namespace octb200 {
/* transcendental forms the reference gets from --use_fast_math (octproz/pri/cuda.pri:54): MUFU approximations */
OCT_HD float oct_lg2(float x) {
#if defined(__CUDA_ARCH__)
	/* lg2.approx.ftz: what the reference's log10f becomes under --use_fast_math (flush-to-zero included); the non-ftz
	 * form costs three more instructions per output for a denormal rescale the reference does not do */
	float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
	return log2f(x);
#endif
}
OCT_HD float oct_sqrt(float x) {
#if defined(__CUDA_ARCH__)
	float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
	return sqrtf(x);
#endif
}
}  // namespace octb200
#define OCT_LG2(x) oct_lg2(x)
#define OCT_SQRT(x) oct_sqrt(x)
#define OCT_FMA(a, b, c) fmaf(a, b, c)

namespace octb200 {

constexpr int XPITCH = 34;                         /* float2 pitch of the 32x32 exchange tile (bank-conflict free for 128-bit reads) */
constexpr int XBUF_FLOAT2 = 32 * XPITCH;           /* per warp */
constexpr int XBUF_BYTES = XBUF_FLOAT2 * 8;

/* stage-A selector */
enum { SA_NONE = 0, SA_LINEAR = 1, SA_LANCZOS = 2, SA_CUBIC = 3 };

struct EpiConsts {
	float scaleA;      /* log: coeff*10*log10(2)/(max-min) applied to log2(p);  lin: coeff/((N/2)*(max-min)) applied to sqrt(p) */
	float scaleB;      /* additive constant (cuda_code.cu:718 / :739 folded on the host in double) */
	int logMode;
	int fpn;           /* subtract meanLine[z] before the magnitude */
	int ppbg;          /* fold postProcessBackgroundRemoval into the store */
	float ppbgWeight, ppbgOffset;
};

OCT_HD int lut_int(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_int(f);
#else
	int i; std::memcpy(&i, &f, 4); return i;
#endif
}

/* ---- stage A.  f points at slot element 0 of the CURRENT line; the float slot has FSLOT_PAD floats in front of the
 * line (plus the Lanczos halo), and for the cubic f[-1] is set to f[1] so that the reference's mirrored first tap
 * abs(n1-1) (cuda_code.cu:284) is a plain n1-1.
 * LUT B[m] = { 4*n1 as int bits (byte offset of tap n1), window*cos(phi), window*sin(phi), t = resample[m] - n1 }.
 * One 16-byte shared-memory read per sample; the interpolation polynomial is evaluated in registers in the
 * reference's form, two samples per packed (f32x2) instruction. */
constexpr int FSLOT_PAD = 4;
/* parity-split float slot of a 2048-sample line: E[k] = f[2k] at float index k, O[k] = f[2k+1] at float index SPLIT_ODD_BASE + k
 * (4 floats of head room before the odd half: O[-1] mirrors f[1] for the cubic's first tap) */
constexpr int SPLIT_ODD_BASE = 1024 + 4;
OCT_HD constexpr bool stage_a_splits_slot(int R, int sa, bool roll) { return R == 2 && !roll && (sa == 3 /* SA_CUBIC */ || sa == 1 /* SA_LINEAR */); }

OCT_HD float ldf(const float* f, int byteOff) {
	return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(f) + byteOff);
}

OCT_HD float2 sample_linear(const float* f, float4 B) {          /* cuda_code.cu:213-231 */
	const int o = lut_int(B.x);
	const float f0 = ldf(f, o), f1 = ldf(f, o + 4);
	const float y = OCT_FMA(f1 - f0, B.w, f0);
	return cscale(make_float2(B.y, B.z), y);
}
/* scalar form (generic pre kernel, odd tails); needs f[-1] == f[1] */
OCT_HD float2 sample_cubic(const float* f, float4 B) {           /* cuda_code.cu:258-295 */
	const int o = lut_int(B.x);
	const float y0 = ldf(f, o - 4), y1 = ldf(f, o), y2 = ldf(f, o + 4), y3 = ldf(f, o + 8);
	const float a = -y0 + 3.0f * (y1 - y2) + y3;
	const float b = 2.0f * y0 - 5.0f * y1 + 4.0f * y2 - y3;
	const float c = -y0 + y2;
	const float pos = B.w, pos2 = pos * pos;
	const float y = 0.5f * pos * (a * pos2 + b * pos + c) + y1;
	return cscale(make_float2(B.y, B.z), y);
}
/* two samples at once, tap-weight form: y = sum_k w_k x[n1-1+k] with the four weights of BOTH samples precomputed per lane
 * (they are the same for every A-scan): linear (cuda_code.cu:229: 0, 1-t, t, 0) or Catmull-Rom (cuda_code.cu:258-271 expanded per
 * tap).  4 packed instructions for two samples instead of 12 for the Horner form.
 * The taps of a sample are f[x], f[x+4], f[y], f[y+4] (byte offsets from the table) with weights W0..W3: in the natural slot
 * y = x + 8 and the weights are in tap order; in the parity-split slot of R = 2 (even samples in one half, odd samples in the
 * other, so that the lanes of a warp -- which walk the line with stride ~1.7 samples -- touch each half with stride < 1 and
 * stay bank-conflict free) x points into the even half, y into the odd half and the host permutes the weights per sample.
 * W[k] = (w of sample a, w of sample b), wa/wb = window*phasor. */
OCT_HD void sample_taps4_x2(const float* f, int xa, int xb, int ya, int yb, float2 W0, float2 W1, float2 W2, float2 W3, float2 wa, float2 wb,
                            float2& outA, float2& outB) {
	const float2 Y0 = make_float2(ldf(f, xa), ldf(f, xb));
	const float2 Y1 = make_float2(ldf(f, xa + 4), ldf(f, xb + 4));
	const float2 Y2 = make_float2(ldf(f, ya), ldf(f, yb));
	const float2 Y3 = make_float2(ldf(f, ya + 4), ldf(f, yb + 4));
	const float2 y = pfma(W3, Y3, pfma(W2, Y2, pfma(W1, Y1, pmul(W0, Y0))));
	/* the two samples are rows j and j + 16 of the lane: their first (twiddle-free) FFT butterfly is folded into the window x
	 * phasor product -- out(j) = wa ya + wb yb, out(j+16) = wa ya - wb yb: 3 packed instructions instead of 4 */
	const float2 t = cscale(wa, y.x);
	outA = pfma(wb, make_float2(y.y, y.y), t);
	outB = pfma(wb, make_float2(-y.y, -y.y), t);
}
/* stage A of these selectors leaves v[j] +- v[j+16] in the registers: the FFT skips its first stage */
OCT_HD constexpr bool stage_a_fuses_stage0(int sa) { return sa == 3 /* SA_CUBIC */ || sa == 1 /* SA_LINEAR */; }

OCT_HD float2 sample_none(const float* f, int m, float4 B) {
	const float y = f[m];
	return cscale(make_float2(B.y, B.z), y);
}

/* cuda_code.cu:297-302 + 304-326: 16 taps i=-7..8 around n0, kernel sinc(pi u) sinc(pi u / 8).
 * shift = 8 for the very first line of the buffer (the reference's offset clamp, cuda_code.cu:313). */
OCT_HD float lanczos8(float x) {
	const float PI_F = 3.141592654f, PI8_F = 0.3926990817f;
	const float ax = fabsf(x);
#if defined(__CUDA_ARCH__)
	const float s1 = __fdividef(__sinf(PI_F * ax), PI_F * ax);
	const float s8 = __fdividef(__sinf(PI8_F * ax), PI8_F * ax);
#else
	const float s1 = sinf(PI_F * ax) / (PI_F * ax);
	const float s8 = sinf(PI8_F * ax) / (PI8_F * ax);
#endif
	return (ax < 0.00001f) ? 1.0f : (s1 * s8);
}

OCT_HD float2 sample_lanczos(const float* f, int shift, float4 B) {
	const int o = lut_int(B.x) + 4 * shift;
	const float t = B.w;
	float sum = 0.0f;
#pragma unroll
	for (int i = -7; i <= 8; ++i) {
		const float y = ldf(f, o + 4 * i);
		sum += y * lanczos8(t - (float)i);
	}
	return cscale(make_float2(B.y, B.z), sum);
}

/* ---- stage A for one lane: 32 samples s = lane + 32 j of sub-sequence p (m = R*s + p) ----
 * LUT ("paired" layout, build_stage_luts_paired): rows j = jj and jj + 16 of a lane share four float4s, stored as four planes
 * of N/2 entries, entry e = p*512 + lane + 32 jj:
 *   P[e] = { wPx_a, wPy_a, wPx_b, wPy_b }   Q[e] = { off_a, off_b, t_a, t_b } (Lanczos / none) or { offX_a, offX_b, offY_a, offY_b } (4-tap)
 *   W01[e] = { w0a, w0b, w1a, w1b }   W23[e] = { w2a, w2b, w3a, w3b }
 * so every packed operand is an aligned register pair straight out of a 128-bit read. */
template <int SA, int R>
OCT_HD void stage_a(int lane, int p, const float* f, int shift, const float4* lut, float2 (&v)[32]) {
	constexpr int HN = 512 * R;
	const float4* P = lut + p * 512;
	const float4* Q = lut + HN + p * 512;
	const float4* W01 = lut + 2 * HN + p * 512;
	const float4* W23 = lut + 3 * HN + p * 512;
#pragma unroll
	for (int jj = 0; jj < 16; ++jj) {
		const float4 Pq = P[lane + 32 * jj];
		const float4 Qq = Q[lane + 32 * jj];
		const float2 wa = make_float2(Pq.x, Pq.y), wb = make_float2(Pq.z, Pq.w);
		if constexpr (SA == SA_CUBIC || SA == SA_LINEAR) {
			const float4 Wa = W01[lane + 32 * jj], Wb = W23[lane + 32 * jj];
			sample_taps4_x2(f, lut_int(Qq.x), lut_int(Qq.y), lut_int(Qq.z), lut_int(Qq.w), make_float2(Wa.x, Wa.y), make_float2(Wa.z, Wa.w),
			                make_float2(Wb.x, Wb.y), make_float2(Wb.z, Wb.w), wa, wb, v[jj], v[jj + 16]);
		} else if constexpr (SA == SA_NONE) {
			const int s = lane + 32 * jj;
			v[jj] = cscale(wa, f[R * s + p]);
			v[jj + 16] = cscale(wb, f[R * (s + 512) + p]);
		} else {
			v[jj] = sample_lanczos(f, shift, make_float4(Qq.x, Pq.x, Pq.y, Qq.z));
			v[jj + 16] = sample_lanczos(f, shift, make_float4(Qq.y, Pq.z, Pq.w, Qq.w));
		}
	}
}

/* ---- four-step exchange: after pass 1, v[r] = V[k1 = bitrev5(r)] for column n2 = lane ----
 * multiply by w_1024^{k1*n2} (tw[k1*32+n2]) and store row-major [k1][n2]. */
OCT_HD void exchange_store(int lane, const float2 (&v)[32], float2* xbuf, const float2* tw) {
	static_for<0, 32>([&](auto rc) {
		constexpr int r = decltype(rc)::value;
		constexpr int k1 = bitrev5(r);
		float2 val = v[r];
		if constexpr (k1 != 0) val = cmul(val, tw[k1 * 32 + lane]);
		xbuf[k1 * XPITCH + lane] = val;
	});
}

/* thread u = lane reads row k1 = u: v[n2] = Y[u][n2] */
OCT_HD void exchange_load(int lane, float2 (&v)[32], const float2* xbuf) {
	const float4* row = reinterpret_cast<const float4*>(xbuf + lane * XPITCH);
#pragma unroll
	for (int q = 0; q < 16; ++q) {
		const float4 t = row[q];
		v[2 * q] = make_float2(t.x, t.y);
		v[2 * q + 1] = make_float2(t.z, t.w);
	}
}

/* ---- R = 2 combine: X[k'] = E0[k'] + w_2048^{k'} E1[k'], k' = lane + 32*k2 < 1024 ----
 * warp 0 finalises k2 < 16, warp 1 finalises k2 >= 16; each hands the other half over through its own tile. */
OCT_HD void combine_store(int lane, int p, float2 (&v)[32], float2* ownTile, const float2* ctw) {
	static_for<0, 32>([&](auto rc) {
		constexpr int r = decltype(rc)::value;
		constexpr int k2 = bitrev5(r);
		if (p == 1) v[r] = cmul(v[r], ctw[lane + 32 * k2]);
		if constexpr (k2 < 16) { if (p == 1) ownTile[k2 * 32 + lane] = v[r]; }
		else                   { if (p == 0) ownTile[(k2 - 16) * 32 + lane] = v[r]; }
	});
}
OCT_HD void combine_load(int lane, int p, float2 (&v)[32], const float2* partnerTile) {
	static_for<0, 32>([&](auto rc) {
		constexpr int r = decltype(rc)::value;
		constexpr int k2 = bitrev5(r);
		if constexpr (k2 < 16) {
			if (p == 0) v[r] = cadd(v[r], partnerTile[k2 * 32 + lane]);
		} else {
			if (p == 1) v[r] = cadd(v[r], partnerTile[(k2 - 16) * 32 + lane]);
		}
	});
}

/* ---- epilogue for 16 outputs per lane: z = zBase + lane + 32*(k2 - K2LO), k2 in [K2LO, K2LO+16) ---- */
OCT_HD float scale_output(float re, float im, const EpiConsts& e) {
	const float pw = OCT_FMA(re, re, im * im);
	return e.logMode ? OCT_FMA(OCT_LG2(pw), e.scaleA, e.scaleB) : OCT_FMA(OCT_SQRT(pw), e.scaleA, e.scaleB);
}
OCT_HD float saturate01(float v) {
#if defined(__CUDA_ARCH__)
	return __saturatef(v);
#else
	if (!(v > 0.0f)) return 0.0f;
	return v > 1.0f ? 1.0f : v;
#endif
}

template <int K2LO, bool LOG, bool FPN, bool PPBG>
OCT_HD void epilogue_scaled_t(int lane, const float2 (&v)[32], const EpiConsts& e, const float2* meanLine,
                              const float* ppbg, float* outLine) {
	const float sA = e.scaleA, sB = e.scaleB, bw = e.ppbgWeight, bo = e.ppbgOffset;
	static_for<0, 32>([&](auto rc) {
		constexpr int r = decltype(rc)::value;
		constexpr int k2 = bitrev5(r);
		if constexpr (k2 >= K2LO && k2 < K2LO + 16) {
			const int z = lane + 32 * k2;
			float2 d = v[r];
			if constexpr (FPN) d = csub(d, meanLine[z]);
			const float pw = OCT_FMA(d.x, d.x, d.y * d.y);
			float o = LOG ? OCT_FMA(OCT_LG2(pw), sA, sB) : OCT_FMA(OCT_SQRT(pw), sA, sB);
			if constexpr (PPBG) o = saturate01(o - OCT_FMA(bw, ppbg[z], bo));
			outLine[z] = o;
		}
	});
}
/* runtime flags -> one uniform branch per line instead of three per output */
template <int K2LO>
OCT_HD void epilogue_scaled(int lane, const float2 (&v)[32], const EpiConsts& e, const float2* meanLine,
                            const float* ppbg, float* outLine) {
	const int sel = (e.logMode ? 1 : 0) | (e.fpn ? 2 : 0) | (e.ppbg ? 4 : 0);
	switch (sel) {
	case 0: epilogue_scaled_t<K2LO, false, false, false>(lane, v, e, meanLine, ppbg, outLine); break;
	case 1: epilogue_scaled_t<K2LO, true, false, false>(lane, v, e, meanLine, ppbg, outLine); break;
	case 2: epilogue_scaled_t<K2LO, false, true, false>(lane, v, e, meanLine, ppbg, outLine); break;
	case 3: epilogue_scaled_t<K2LO, true, true, false>(lane, v, e, meanLine, ppbg, outLine); break;
	case 4: epilogue_scaled_t<K2LO, false, false, true>(lane, v, e, meanLine, ppbg, outLine); break;
	case 5: epilogue_scaled_t<K2LO, true, false, true>(lane, v, e, meanLine, ppbg, outLine); break;
	case 6: epilogue_scaled_t<K2LO, false, true, true>(lane, v, e, meanLine, ppbg, outLine); break;
	default: epilogue_scaled_t<K2LO, true, true, true>(lane, v, e, meanLine, ppbg, outLine); break;
	}
}

/* complex store of the same 16 bins (fixed-pattern-noise determination pre-pass) */
template <int K2LO>
OCT_HD void epilogue_complex(int lane, const float2 (&v)[32], float2* outLine) {
	static_for<0, 32>([&](auto rc) {
		constexpr int r = decltype(rc)::value;
		constexpr int k2 = bitrev5(r);
		if constexpr (k2 >= K2LO && k2 < K2LO + 16) outLine[lane + 32 * k2] = v[r];
	});
}

}  // namespace octb200
