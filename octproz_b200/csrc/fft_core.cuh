/*
 * fft_core.cuh -- in-register 32-point inverse DFT network and the 32x32 four-step index maps
 * used by the fused OCT kernels.  Replaces cufftExecC2C(..., CUFFT_INVERSE) of the reference
 * (cuda_code.cu:1140,1514-1515): X[k] = sum_m x[m] exp(+2 pi i m k / N), unnormalised.
 *
 * The functions are __host__ __device__ so tests/emu can execute the exact lane/register maps on
 * the CPU (test-only emulator; the product never runs this code on the host).
 */
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define OCT_HD __host__ __device__ __forceinline__
#else
#include <vector_types.h>
#include <vector_functions.h>
#define OCT_HD inline
#endif

#include <type_traits>
#include <cmath>

namespace octb200 {

template <int B, int E, class F>
OCT_HD void static_for(F&& f) {
	if constexpr (B < E) {
		f(std::integral_constant<int, B>{});
		static_for<B + 1, E>(f);
	}
}

OCT_HD constexpr int bitrev5(int r) {
	return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4);
}

/* w_32^k = exp(+2 pi i k / 32), k = 0..15 */
template <int K> struct W32 {
	static constexpr float c[16] = {
		1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
		0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
		0.0f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
		-0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f };
	static constexpr float s[16] = {
		0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
		0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
		1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
		0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f };
	static constexpr float re = c[K];
	static constexpr float im = s[K];
};

/* ---- complex arithmetic on packed fp32 pairs ----
 * sm_100a has two-lane fp32 instructions (PTX add/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2) whose operands take
 * per-half swizzles (LO_HI), per-half negation and scalar broadcast for free.  A complex value is exactly one such
 * pair, so a butterfly with a general twiddle is 4 issue slots instead of 8.  Host builds (tests/emu) use plain floats. */
OCT_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	return __fadd2_rn(a, b);
#else
	return make_float2(a.x + b.x, a.y + b.y);
#endif
}
OCT_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
#else
	return make_float2(a.x - b.x, a.y - b.y);
#endif
}
/* a * (s, s) */
OCT_HD float2 cscale(float2 a, float s) {
#if defined(__CUDA_ARCH__)
	return __fmul2_rn(a, make_float2(s, s));
#else
	return make_float2(a.x * s, a.y * s);
#endif
}
/* generic packed pair helpers (two independent fp32 lanes) */
OCT_HD float2 pfma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
	return __ffma2_rn(a, b, c);
#else
	return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
OCT_HD float2 pmul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	return __fmul2_rn(a, b);
#else
	return make_float2(a.x * b.x, a.y * b.y);
#endif
}
OCT_HD float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }

/* i * a = (-a.y, a.x): a swizzle + half negation, folded into the consumer's operand modifiers */
OCT_HD float2 cmul_i(float2 a) { return make_float2(-a.y, a.x); }

/* c = a * b (complex): 2 packed instructions */
OCT_HD float2 cmul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
	const float2 t = __fmul2_rn(make_float2(a.x, a.x), b);
	return __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), t);
#else
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}

/* d * w_32^IDX with the trivial cases folded at compile time */
template <int IDX>
OCT_HD float2 mul_w32(float2 d) {
	if constexpr (IDX == 0) {
		return d;
	} else if constexpr (IDX == 8) {
		return cmul_i(d);
	} else if constexpr (IDX == 4) {
		constexpr float h = 0.70710678118654752440f;
		return cscale(cadd(d, cmul_i(d)), h);             /* (x - y, x + y) h */
	} else if constexpr (IDX == 12) {
		constexpr float h = 0.70710678118654752440f;
		return cscale(csub(cmul_i(d), d), h);             /* (-x - y, x - y) h */
	} else {
		constexpr float c = W32<IDX>::re, s = W32<IDX>::im;
#if defined(__CUDA_ARCH__)
		const float2 t = __fmul2_rn(d, make_float2(c, c));
		return __ffma2_rn(make_float2(d.y, d.x), make_float2(-s, s), t);
#else
		return make_float2(d.x * c - d.y * s, d.x * s + d.y * c);
#endif
	}
}

/*
 * In-place radix-2 decimation-in-frequency inverse DFT of 32 complex registers.
 * Input  v[j]           = x[j]
 * Output v[r]           = X[bitrev5(r)]      (bit-reversed register order, all indices compile time)
 */
OCT_HD void fft32_inv_dif(float2 (&v)[32]) {
	static_for<0, 5>([&](auto sc) {
		constexpr int s = decltype(sc)::value;
		constexpr int half = 16 >> s;
		static_for<0, (1 << s)>([&](auto gc) {
			constexpr int g = decltype(gc)::value;
			static_for<0, half>([&](auto kc) {
				constexpr int k = decltype(kc)::value;
				constexpr int i0 = g * 2 * half + k;
				constexpr int i1 = i0 + half;
				const float2 a = v[i0], b = v[i1];
				v[i0] = cadd(a, b);
				v[i1] = mul_w32<(k << s)>(csub(a, b));
			});
		});
	});
}

/* tan(2 pi k / 32), k = 0..15 (k = 8 is never used: w = i is handled as a swizzle) */
template <int K> struct T32 {
	static constexpr float t[16] = {
		0.0f, 0.19891236737965800691f, 0.41421356237309504880f, 0.66817863791929891999f,
		1.0f, 1.49660576266548901380f, 2.41421356237309504880f, 5.02733949212584810451f,
		0.0f, -5.02733949212584810451f, -2.41421356237309504880f, -1.49660576266548901380f,
		-1.0f, -0.66817863791929891999f, -0.41421356237309504880f, -0.19891236737965800691f };
	static constexpr float v = t[K];
};

/*
 * Decimation-in-time butterfly with the twiddle folded into fused multiply-adds (Linzer-Feig form):
 *   w b = cos * (b + tan * i b),   a' = a + w b,   b' = a - w b
 * is 3 packed FFMA2 (6 scalar FMAs) instead of the 4 packed instructions of "multiply, then add/subtract".
 * The rounding error of the scaled form is bounded by eps |b| because cos * tan = sin <= 1.
 */
template <int IDX>
OCT_HD void bf_dit(float2& a, float2& b) {
	if constexpr (IDX == 0) {
		const float2 x = a, y = b;
		a = cadd(x, y);
		b = csub(x, y);
	} else if constexpr (IDX == 8) {
		const float2 x = a, y = cmul_i(b);
		a = cadd(x, y);
		b = csub(x, y);
	} else if constexpr (IDX == 4 || IDX == 12) {
		constexpr float h = 0.70710678118654752440f;
		const float2 x = a;
		const float2 u = (IDX == 4) ? cadd(b, cmul_i(b)) : csub(cmul_i(b), b);     /* (1 + i) b  or  (-1 + i) b */
		a = pfma(u, make_float2(h, h), x);
		b = pfma(u, make_float2(-h, -h), x);
	} else {
		constexpr float c = W32<IDX>::re, t = T32<IDX>::v;
		const float2 x = a;
		const float2 u = pfma(make_float2(t, t), cmul_i(b), b);
		a = pfma(u, make_float2(c, c), x);
		b = pfma(u, make_float2(-c, -c), x);
	}
}

constexpr int bitrev_n(int g, int bits) {
	int r = 0;
	for (int i = 0; i < bits; ++i) r |= ((g >> i) & 1) << (bits - 1 - i);
	return r;
}

/*
 * In-place radix-2 decimation-in-time inverse DFT of 32 complex registers, natural-order input:
 * stage s pairs the registers that differ in index bit 4 - s; the twiddle w_32^{(16 >> s) rev_s(g)} depends only on the
 * already transformed top s index bits g and multiplies the upper-index element before the butterfly.
 * Input  v[j] = x[j],  output v[r] = X[bitrev5(r)]  (same maps as fft32_inv_dif).
 * Unused outputs (the pruned upper half of the spectrum) are removed by dead-code elimination.
 */
template <bool SKIP0 = false>
OCT_HD void fft32_inv_dit(float2 (&v)[32]) {
	/* SKIP0: the caller has already formed v[j] +- v[j + 16] (stage A folds that butterfly into its window x phasor product) */
	static_for<(SKIP0 ? 1 : 0), 5>([&](auto sc) {
		constexpr int s = decltype(sc)::value;
		constexpr int half = 16 >> s;
		static_for<0, (1 << s)>([&](auto gc) {
			constexpr int g = decltype(gc)::value;
			constexpr int idx = (16 >> s) * bitrev_n(g, s);
			static_for<0, half>([&](auto kc) {
				constexpr int k = decltype(kc)::value;
				constexpr int i0 = g * 2 * half + k;
				bf_dit<idx>(v[i0], v[i0 + half]);
			});
		});
	});
}

/* the network the kernels use (OCT_FFT_DIF selects the plain multiply-then-butterfly form for A/B builds) */
template <bool SKIP0 = false>
OCT_HD void fft32_inv(float2 (&v)[32]) {
#ifdef OCT_FFT_DIF
	if constexpr (SKIP0) {      /* A/B builds: redo nothing, the first DIF stage has a twiddle -- not supported together */
		static_assert(!SKIP0, "OCT_FFT_DIF cannot skip stage 0");
	}
	fft32_inv_dif(v);
#else
	fft32_inv_dit<SKIP0>(v);
#endif
}

}  // namespace octb200
