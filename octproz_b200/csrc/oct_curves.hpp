/*
 * oct_curves.hpp -- HOST curve generators of the product (header-only): what
 * OctAlgorithmParameters::update{Resample,Dispersion,Window}Curve produce in the reference
 * (octalgorithmparameters.cpp:141-249 -> polynomial.cpp:108-145, windowfunction.cpp:121-253) and what
 * fillSinusoidalScanCorrectionCurve produces on the GPU (cuda_code.cu:516-521).
 *
 * Written from the reference's arithmetic, not from its code: results are bit-identical fp32 LUTs
 * (pinned by tests/test_curves.py against tests/golden/luts_*.npz, generated from the reference's own
 * sources compiled in place).  Build with -ffp-contract=off (the reference's host build has no FMA
 * contraction; the one explicit fma is the Horner step).
 */
#pragma once
#include <cmath>
#include <vector>

namespace octb200 {
namespace curves {

/* third-order polynomial with the reference's coefficient normalisation: k_i / (N-1)^i in fp32,
 * evaluated by float-FMA Horner at x = 0..N-1 */
inline void cubic_polynomial(int n, float k0, float k1, float k2, float k3, float* out) {
	const float d = static_cast<float>(n - 1);
	const float c0 = k0;
	const float c1 = k1 / d;
	const float c2 = k2 / std::pow(d, 2.0f);   /* powf */
	const float c3 = k3 / std::pow(d, 3.0f);
	for (int i = 0; i < n; ++i) {
		const float x = static_cast<float>(i);
		float r = std::fma(0.0f, x, c3);
		r = std::fma(r, x, c2);
		r = std::fma(r, x, c1);
		r = std::fma(r, x, c0);
		out[i] = r;
	}
}

/* values outside [0, N-3] would make the interpolation kernels read outside the A-scan */
inline void clamp_resample(int n, float* curve) {
	const float lo = 0.0f, hi = static_cast<float>(static_cast<unsigned>(n - 3));
	for (int i = 0; i < n; ++i) {
		if (curve[i] < lo) curve[i] = lo;
		if (curve[i] > hi) curve[i] = hi;
	}
}

inline void resample(int n, float c0, float c1, float c2, float c3, float* out) {
	cubic_polynomial(n, c0, c1, c2, c3, out);
	clamp_resample(n, out);
}
inline void dispersion(int n, float d0, float d1, float d2, float d3, float* out) {
	cubic_polynomial(n, d0, d1, d2, d3, out);
}

enum Window { Hanning = 0, Gauss = 1, Sine = 2, Lanczos = 3, Rectangular = 4, FlatTop = 5 };

inline void window(int type, float center, float fill, int n, float* out) {
	constexpr double kPi = 3.14159265358979323846;
	if (center > 1.0f) center = 1.0f;
	else if (center < 0.0f) center = 0.0f;
	const unsigned size = static_cast<unsigned>(n);
	if (type == Gauss) {
		const unsigned c = static_cast<unsigned>(center * size);
		for (unsigned i = 0; i < size; ++i) {
			const int xi = static_cast<int>(i) - static_cast<int>(c);
			const float xn = (static_cast<float>(xi) / (static_cast<float>(size) - 1.0f)) / fill;
			out[i] = std::exp(-10.0f * std::pow(xn, 2.0f));   /* expf / powf */
		}
		return;
	}
	/* support of the window: `fill` of the line, centred at `center`; unsigned arithmetic as in the reference */
	const unsigned width = static_cast<unsigned>(fill * size);
	const unsigned c = static_cast<unsigned>(center * size);
	int first = static_cast<int>(c - width / 2);
	if (first + static_cast<int>(width) < first) first = first + static_cast<int>(width);
	for (unsigned i = 0; i < size; ++i) {
		const float xn = static_cast<float>(static_cast<int>(i) - first) / (static_cast<float>(width) - 1.0f);
		float v = 0.0f;
		if (!(xn > 0.999f || xn < 0.0001f)) {
			const double x = static_cast<double>(xn);
			switch (type) {
			case Hanning: v = static_cast<float>(0.5 * (1.0 - std::cos(2.0 * kPi * x))); break;
			case Sine:    v = static_cast<float>(std::sin(kPi * x)); break;
			case Lanczos: {
				const float a = 2.0f * xn - 1.0f;
				v = (a == 0.0f) ? 1.0f : static_cast<float>(std::sin(kPi * static_cast<double>(a)) / (kPi * static_cast<double>(a)));
				break;
			}
			case FlatTop: {
				const float a0 = 0.215578948f, a1 = 0.416631580f, a2 = 0.277263158f, a3 = 0.083578947f, a4 = 0.006947368f;
				v = a0 - a1 * static_cast<float>(std::cos(2.0 * kPi * x)) + a2 * static_cast<float>(std::cos(4.0 * kPi * x))
				       - a3 * static_cast<float>(std::cos(6.0 * kPi * x)) + a4 * static_cast<float>(std::cos(8.0 * kPi * x));
				break;
			}
			case Rectangular:
			default: v = 1.0f; break;
			}
		}
		out[i] = v;
	}
}

/* curve[k] = (A/pi) * acos(1 - 2k/A): position of lateral sample k of a sinusoidal fast-axis scan */
inline void sinusoidal(int ascans, float* out) {
	constexpr double kPi = 3.14159265358979323846;
	const double len = static_cast<double>(static_cast<float>(ascans));
	for (int k = 0; k < ascans; ++k) {
		const float arg = static_cast<float>(1.0 - (2.0 * static_cast<double>(static_cast<float>(k))) / len);
		out[k] = static_cast<float>((len / kPi) * std::acos(static_cast<double>(arg)));
	}
}

}  // namespace curves
}  // namespace octb200
