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

}  // namespace octb200
