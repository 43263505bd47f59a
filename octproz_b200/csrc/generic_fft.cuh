/*
 * generic_fft.cuh -- the mixed-radix inverse transform of the shared-memory kernel (k_generic.cu): butterflies, one Stockham autosort
 * pass, the padded line-buffer index map, and the host-side plan / twiddle-table builders.  Replaces cufftPlan1d / cufftExecC2C
 * (CUFFT_INVERSE) of the reference (cuda_code.cu:1140,1514-1515) for line lengths other than 1024 / 2048.
 *
 * The functions are __host__ __device__ so that tests/emu can execute the exact pass / index arithmetic on the CPU for every
 * supported length (test-only emulator; the product never runs this code on the host).
 */
#pragma once
#include "fft_core.cuh"

#include <cmath>

namespace octb200 {

#if defined(__CUDACC__)
#define OCT_UNROLL _Pragma("unroll")
#else
#define OCT_UNROLL
#endif

OCT_HD unsigned mulhi_u32(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (unsigned)(((unsigned long long)a * (unsigned long long)b) >> 32);
#endif
}

template <int P> struct DftTab;
#include "k_generic_tables.inc"

/* ---- butterflies: inverse DFT of R register values, X_j = sum_k x_k exp(+2 pi i j k / R) ---- */
template <int R>
OCT_HD void dft_inv(float2 (&v)[R]) {
	if constexpr (R == 2) {
		const float2 a = v[0], b = v[1];
		v[0] = cadd(a, b); v[1] = csub(a, b);
	} else if constexpr (R == 4) {
		const float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
		const float2 s13 = cadd(v[1], v[3]), d13 = cmul_i(csub(v[1], v[3]));        /* +i (v1 - v3) */
		v[0] = cadd(s02, s13); v[2] = csub(s02, s13);
		v[1] = cadd(d02, d13); v[3] = csub(d02, d13);
	} else if constexpr (R == 8) {
		float2 e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
		dft_inv<4>(e); dft_inv<4>(o);
		const float2 t1 = mul_w32<4>(o[1]), t2 = mul_w32<8>(o[2]), t3 = mul_w32<12>(o[3]);      /* w_8^k = w_32^{4k} */
		v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
		v[1] = cadd(e[1], t1);   v[5] = csub(e[1], t1);
		v[2] = cadd(e[2], t2);   v[6] = csub(e[2], t2);
		v[3] = cadd(e[3], t3);   v[7] = csub(e[3], t3);
	} else {
		/* odd prime: pair k with R - k.  a_k = x_k + x_{R-k}, b_k = x_k - x_{R-k};
		 * X_j = x_0 + sum a_k cos(2 pi j k / R) + i sum b_k sin(2 pi j k / R), X_{R-j} the same with - i */
		constexpr int H = (R - 1) / 2;
		float2 a[H], b[H];
		static_for<0, H>([&](auto kc) {
			constexpr int k = decltype(kc)::value + 1;
			a[k - 1] = cadd(v[k], v[R - k]);
			b[k - 1] = csub(v[k], v[R - k]);
		});
		const float2 x0 = v[0];
		float2 sum = x0;
		static_for<0, H>([&](auto kc) { sum = cadd(sum, a[decltype(kc)::value]); });
		v[0] = sum;
		static_for<0, H>([&](auto jc) {
			constexpr int j = decltype(jc)::value + 1;
			float2 C = x0, S = make_float2(0.f, 0.f);
			static_for<0, H>([&](auto kc) {
				constexpr int k = decltype(kc)::value + 1;
				constexpr float c = DftTab<R>::c[(j * k) % R], s = DftTab<R>::s[(j * k) % R];
				C = pfma(a[k - 1], make_float2(c, c), C);
				S = pfma(b[k - 1], make_float2(s, s), S);
			});
			const float2 iS = cmul_i(S);
			v[j] = cadd(C, iS);
			v[R - j] = csub(C, iS);
		});
	}
}

/* line buffers are padded by one element per 32 so that the strided accesses of the first passes (stride R elements) spread over the banks */
OCT_HD int gpad(int i) { return i + (i >> 5); }
OCT_HD int gpad_len(int N) { return N + (N >> 5) + 1; }

/* one Stockham autosort pass of radix R over a line of N complex values in shared memory (Ns = product of the earlier radices):
 * butterfly j reads in[j + i N/R], multiplies by w_{Ns R}^{i k}, k = j mod Ns, transforms, and writes out[(j - k) R + k + i Ns].
 * tw = this pass's table (shared memory), tw[(i - 1) Ns + k] = w_{Ns R}^{i k}: the lanes of a warp (consecutive k) read consecutive words.
 * j mod Ns by multiplication: magic = ceil(2^32 / Ns) is exact for j Ns < 2^32 / Ns, i.e. for every N <= 8192 */
template <int R>
OCT_HD void stockham_pass(const float2* __restrict__ in, float2* __restrict__ out, int N, int Ns, unsigned magic,
                                              const float2* __restrict__ tw, int tid, int T) {
	const int M = N / R;
	for (int j = tid; j < M; j += T) {
		float2 v[R];
		OCT_UNROLL
		for (int i = 0; i < R; ++i) v[i] = in[gpad(j + i * M)];
		int k = 0;
		if (Ns > 1) {
			k = j - (int)mulhi_u32((unsigned)j, magic) * Ns;
			OCT_UNROLL
			for (int i = 1; i < R; ++i) v[i] = cmul(v[i], tw[(i - 1) * Ns + k]);
		}
		dft_inv<R>(v);
		const int j0 = (j - k) * R + k;
		OCT_UNROLL
		for (int i = 0; i < R; ++i) out[gpad(j0 + i * Ns)] = v[i];
	}
}

/* radix plan: odd primes first, then the power of two as 8s with the remainder as 4 (2^{3q+2}), 4 * 4 (2^{3q+1}, q >= 1) or 2 */
inline bool generic_fft_plan(int N, int* radix, int* nPass) {
	if (N < 8 || (N & 1) || N > 8192) return false;
	int n = N, cnt = 0;
	const int odd[] = { 13, 11, 7, 5, 3 };
	for (int p : odd) while (n % p == 0) { if (cnt >= 12) return false; radix[cnt++] = p; n /= p; }
	int a = 0;
	while (n % 2 == 0) { ++a; n /= 2; }
	if (n != 1) return false;
	int eights = a / 3;
	const int rem = a % 3;
	if (rem == 1 && eights >= 1) { --eights; for (int i = 0; i < eights; ++i) radix[cnt++] = 8; radix[cnt++] = 4; radix[cnt++] = 4; }
	else { for (int i = 0; i < eights; ++i) radix[cnt++] = 8; if (rem == 2) radix[cnt++] = 4; else if (rem == 1) radix[cnt++] = 2; }
	*nPass = cnt;
	return cnt <= 16;
}

/* per-pass twiddle tables, concatenated: pass p (radix R, Ns = product of the earlier radices) holds tw[(i - 1) Ns + k] =
 * exp(+2 pi i * i k / (Ns R)), i = 1 .. R-1, k < Ns; the first pass (Ns = 1) has none.  Returns the total number of entries. */
inline int generic_twiddle_layout(int N, const int* radix, int nPass, int* twOff, unsigned* magic) {
	(void)N;
	int off = 0, Ns = 1;
	for (int p = 0; p < nPass; ++p) {
		twOff[p] = off;
		magic[p] = (Ns > 1) ? (unsigned)((0x100000000ull + (unsigned long long)Ns - 1ull) / (unsigned long long)Ns) : 0u;
		if (Ns > 1) off += (radix[p] - 1) * Ns;
		Ns *= radix[p];
	}
	return off > 0 ? off : 1;
}
inline void generic_fill_twiddles(const int* radix, int nPass, const int* twOff, float2* tw) {
	int Ns = 1;
	for (int p = 0; p < nPass; ++p) {
		const int R = radix[p];
		if (Ns > 1)
			for (int i = 1; i < R; ++i)
				for (int k = 0; k < Ns; ++k) {
					const double ang = 2.0 * 3.14159265358979323846 * (double)i * (double)k / ((double)Ns * (double)R);
					tw[twOff[p] + (i - 1) * Ns + k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
				}
		Ns *= R;
	}
}

}  // namespace octb200
