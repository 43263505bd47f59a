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

/* d * w_32^IDX with the trivial cases folded at compile time */
template <int IDX>
OCT_HD float2 mul_w32(float2 d) {
	if constexpr (IDX == 0) {
		return d;
	} else if constexpr (IDX == 8) {
		return make_float2(-d.y, d.x);
	} else if constexpr (IDX == 4) {
		constexpr float h = 0.70710678118654752440f;
		return make_float2((d.x - d.y) * h, (d.x + d.y) * h);
	} else if constexpr (IDX == 12) {
		constexpr float h = 0.70710678118654752440f;
		return make_float2((-d.x - d.y) * h, (d.x - d.y) * h);
	} else {
		constexpr float c = W32<IDX>::re, s = W32<IDX>::im;
		return make_float2(d.x * c - d.y * s, d.x * s + d.y * c);
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
				v[i0] = make_float2(a.x + b.x, a.y + b.y);
				v[i1] = mul_w32<(k << s)>(make_float2(a.x - b.x, a.y - b.y));
			});
		});
	});
}

/* c = a * b (complex) */
OCT_HD float2 cmul(float2 a, float2 b) {
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

}  // namespace octb200
