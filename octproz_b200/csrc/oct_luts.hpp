/*
 * oct_luts.hpp -- HOST-side table builders for the fused kernels (header-only, plain C++).
 *
 * The reference uploads three per-sample curves (resample, window, dispersion phase) and lets every
 * thread re-derive tap indices / interpolation polynomials / the window*phasor product for every
 * A-scan (cuda_code.cu:213-489).  All of that is identical for every A-scan, so it is folded ONCE into
 * one float4 table entry per sample m:
 *   B[m] = { 4*(int)resample[m] (byte offset), window*cos(phi), window*sin(phi), resample[m] - (int)resample[m] }
 * plus the four-step twiddles of the 32x32 FFT and the folded scale constants of
 * postProcessTruncateLog/Lin (cuda_code.cu:699-741).
 */
#pragma once
#include "oct_phases.cuh"

#include <cmath>
#include <cstring>
#include <vector>

namespace octb200 {

struct StageLuts {
	std::vector<float4> B;
};

inline float int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

/* entry of sample m lives at (m % R) * (N / R) + m / R so that each of the R warps of a line group
 * reads consecutive 16-byte words (bank-conflict free) */
inline size_t lut_slot(int m, int N, int R) { return (size_t)(m % R) * (size_t)(N / R) + (size_t)(m / R); }

/*
 * N           samples per line
 * R           warps per line group (1 for N=1024, 2 for N=2048); 1 for the generic pre-FFT kernel
 * resample    clamped resample curve (nullptr when resampling is off)
 * window      window LUT or nullptr (windowing off -> 1)
 * phasor      (cos phi, sin phi) pairs as produced by fill_phase on the GPU, or nullptr (dispersion off -> (1,0))
 */
inline void build_stage_luts(int N, int R, const float* resample, const float* window, const float2* phasor, StageLuts& out) {
	out.B.assign((size_t)N, make_float4(0, 0, 0, 0));
	for (int m = 0; m < N; ++m) {
		const float w = window ? window[m] : 1.0f;
		const float px = phasor ? phasor[m].x : 1.0f;
		const float py = phasor ? phasor[m].y : 0.0f;
		const float xi = resample ? resample[m] : (float)m;
		const int n1 = (int)xi;                 /* C truncation, cuda_code.cu:223 / :283 / :315 */
		const float t = xi - (float)n1;         /* exact in fp32 */
		out.B[lut_slot(m, N, R)] = make_float4(int_as_float(4 * n1), w * px, w * py, t);   /* byte offset of tap n1 */
	}
}

/* tap weights of the 4-tap interpolators for fractional position t (double precision, rounded once):
 * interp 1 = Catmull-Rom, the per-tap expansion of cuda_code.cu:263-270; otherwise linear (cuda_code.cu:229) */
inline void tap_weights(int interp, float tf, float (&w)[4]) {
	const double t = (double)tf;
	if (interp == 1) {
		w[0] = (float)(0.5 * (-t * t * t + 2.0 * t * t - t));
		w[1] = (float)(0.5 * (3.0 * t * t * t - 5.0 * t * t + 2.0));
		w[2] = (float)(0.5 * (-3.0 * t * t * t + 4.0 * t * t + t));
		w[3] = (float)(0.5 * (t * t * t - t * t));
	} else {
		w[0] = 0.0f; w[1] = (float)(1.0 - t); w[2] = (float)t; w[3] = 0.0f;
	}
}

/* natural-order 4-tap table of the shared-memory kernel (k_generic.cu): two float4 per sample,
 * { byte offset of tap n1 - 1, w0, w1, w2 } { w3, window * cos, window * sin, 0 }; f[-1] of the float slot mirrors f[1] (cuda_code.cu:284) */
inline void build_stage_luts_taps(int N, int interp, const float* resample, const float* window, const float2* phasor, std::vector<float4>& out) {
	StageLuts nat;
	build_stage_luts(N, 1, resample, window, phasor, nat);
	out.assign((size_t)2 * N, make_float4(0, 0, 0, 0));
	for (int m = 0; m < N; ++m) {
		const float4 B = nat.B[m];
		int o; std::memcpy(&o, &B.x, 4);
		float w[4];
		tap_weights(interp, B.w, w);
		out[2 * m] = make_float4(int_as_float(o - 4), w[0], w[1], w[2]);
		out[2 * m + 1] = make_float4(w[3], B.y, B.z, 0.0f);
	}
}

/* paired layout for the fused kernel (see stage_a): 2N float4s = four planes [ P | Q | W01 | W23 ] of N/2 entries.
 * taps = 0: Q = { off_a, off_b, t_a, t_b } (byte offset of tap n1 and fractional position: Lanczos, no resampling)
 * taps = 1: 4-tap interpolators, natural float slot: Q = { offX_a, offX_b, offY_a, offY_b } with offX = 4 (n1 - 1), offY = offX + 8,
 *           weights in tap order
 * taps = 2: 4-tap interpolators, parity-split slot (R = 2): the taps n1-1 .. n1+2 are two consecutive elements of the even half (X) and
 *           two of the odd half (Y); the weights are permuted per sample to match */
inline void build_stage_luts_paired(int N, int R, int interp, const float* resample, const float* window, const float2* phasor,
                                    std::vector<float4>& out, int taps = 0) {
	StageLuts nat;
	build_stage_luts(N, 1, resample, window, phasor, nat);      /* natural order: nat.B[m] = {off, wPx, wPy, t} */
	out.assign((size_t)2 * N, make_float4(0, 0, 0, 0));
	const int half = N / 2;
	auto tap_offsets = [&](float offBits, float& ox, float& oy, float (&w)[4]) {
		int o; std::memcpy(&o, &offBits, 4);
		const int t0 = o / 4 - 1;                               /* first tap n1 - 1 (-1 = the mirrored pad sample) */
		if (taps == 1) { ox = int_as_float(4 * t0); oy = int_as_float(4 * t0 + 8); return; }
		/* X is ALWAYS the even half and Y the odd half: one gather instruction of a warp then stays inside one half (lane stride
		 * ~0.85 floats, conflict free); mixing the halves in one instruction collides them bank against bank */
		const int k = (t0 >= 0) ? t0 / 2 : -1;                  /* floor(t0 / 2) for t0 >= -1 */
		const float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
		oy = int_as_float(4 * (SPLIT_ODD_BASE + k));            /* O[k] O[k+1] */
		if ((t0 & 1) == 0) { ox = int_as_float(4 * k);       w[0] = w0; w[1] = w2; w[2] = w1; w[3] = w3; }    /* taps E[k] O[k] E[k+1] O[k+1] */
		else               { ox = int_as_float(4 * (k + 1)); w[0] = w1; w[1] = w3; w[2] = w0; w[3] = w2; }    /* taps O[k] E[k+1] O[k+1] E[k+2] */
	};
	for (int p = 0; p < R; ++p)
		for (int jj = 0; jj < 16; ++jj)
			for (int lane = 0; lane < 32; ++lane) {
				const int sa = lane + 32 * jj, sb = sa + 512;              /* rows j = jj and jj + 16: the two inputs of the first FFT butterfly */
				const float4 A = nat.B[(size_t)R * sa + p], B = nat.B[(size_t)R * sb + p];
				float wa[4], wb[4];
				tap_weights(interp, A.w, wa); tap_weights(interp, B.w, wb);
				const size_t idx = (size_t)p * 512 + lane + 32 * jj;
				out[idx] = make_float4(A.y, A.z, B.y, B.z);
				if (taps == 0) {
					out[half + idx] = make_float4(A.x, B.x, A.w, B.w);
				} else {
					float xa, ya, xb, yb;
					tap_offsets(A.x, xa, ya, wa); tap_offsets(B.x, xb, yb, wb);
					out[half + idx] = make_float4(xa, xb, ya, yb);
				}
				out[2 * half + idx] = make_float4(wa[0], wb[0], wa[1], wb[1]);
				out[3 * half + idx] = make_float4(wa[2], wb[2], wa[3], wb[3]);
			}
}

/* tw[k1*32+n2] = exp(+2 pi i k1 n2 / 1024): inter-pass twiddles of the 32x32 four-step transform */
inline void build_twiddles_1024(std::vector<float2>& tw) {
	tw.resize(1024);
	for (int k1 = 0; k1 < 32; ++k1)
		for (int n2 = 0; n2 < 32; ++n2) {
			const double a = 2.0 * M_PI * (double)(k1 * n2) / 1024.0;
			tw[k1 * 32 + n2] = make_float2((float)std::cos(a), (float)std::sin(a));
		}
}
/* ctw[k] = exp(+2 pi i k / 2048), k < 1024: radix-2 combine of the two interleaved 1024-point transforms */
inline void build_combine_twiddles_2048(std::vector<float2>& ctw) {
	ctw.resize(1024);
	for (int k = 0; k < 1024; ++k) {
		const double a = 2.0 * M_PI * (double)k / 2048.0;
		ctw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
	}
}

/* cuda_code.cu:718 and :739 with all constants folded in double */
inline EpiConsts make_epi_consts(int N, int logMode, float gmin, float gmax, float coeff, float addend) {
	EpiConsts e;
	std::memset(&e, 0, sizeof(e));
	const double h = (double)(N / 2);
	const double range = (double)gmax - (double)gmin;
	e.logMode = logMode;
	if (logMode) {
		e.scaleA = (float)((double)coeff * 10.0 * std::log10(2.0) / range);
		e.scaleB = (float)((double)coeff * ((-10.0 * std::log10(h) - (double)gmin) / range + (double)addend));
	} else {
		e.scaleA = (float)((double)coeff / (h * range));
		e.scaleB = (float)((double)coeff * (-(double)gmin / range + (double)addend));
	}
	return e;
}

}  // namespace octb200
