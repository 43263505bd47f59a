/*
 * oct_oracle.c -- CPU restatement of the OCTproZ raw->B-scan arithmetic.
 * TEST INFRASTRUCTURE ONLY -- see oct_oracle.h for the rules and the parity status.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  -ffp-contract=off
 * matters: the reference's host code is built without FMA contraction (qmake -O2, SSE2).
 *
 * Paths are relative to /root/reference/octproz_project/octproz/src/ (CU = cuda_code.cu).
 */
#define _GNU_SOURCE
#include "oct_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ LUTs */

/* polynomial.cpp:108-116 getValueAt: float-FMA Horner from the highest coefficient */
static float poly_value(const float* c, int order, float x) {
	float r = 0.0f;
	for (int i = 0; i <= order; ++i) r = fmaf(r, x, c[order - i]);
	return r;
}

/* octalgorithmparameters.cpp:149-156 coefficient scaling (float / powf) */
static void scaled_coeffs(int N, float k0, float k1, float k2, float k3, float* c) {
	c[0] = k0;
	c[1] = k1 / (float)(N - 1);
	c[2] = k2 / powf((float)(N - 1), 2);
	c[3] = k3 / powf((float)(N - 1), 3);
}

void orc_resample_curve(int N, float c0, float c1, float c2, float c3, float* out) {
	float c[4];
	scaled_coeffs(N, c0, c1, c2, c3, c);
	for (int i = 0; i < N; ++i) out[i] = poly_value(c, 3, (float)i);   /* polynomial.cpp:139-145 */
	/* polynomial.cpp:126-137 clamp to [0, N-3]; octalgorithmparameters.cpp:167 */
	const float lo = 0.0f, hi = (float)(unsigned)(N - 3);
	for (int i = 0; i < N; ++i) {
		if (out[i] < lo) out[i] = lo;
		if (out[i] > hi) out[i] = hi;
	}
}

void orc_dispersion_curve(int N, float d0, float d1, float d2, float d3, float* out) {
	float c[4];
	scaled_coeffs(N, d0, d1, d2, d3, c);                                /* octalgorithmparameters.cpp:210-213 */
	for (int i = 0; i < N; ++i) out[i] = poly_value(c, 3, (float)i);
}

/* windowfunction.cpp:121-253 */
void orc_window_curve(int type, float center, float fill, int N, float* out) {
	/* windowfunction.cpp:65-73 centre position clamped to [0,1] */
	if (center > 1) center = 1.0f; else if (center < 0) center = 0;
	const unsigned size = (unsigned)N;
	if (type == ORC_WIN_GAUSS) {                                        /* :164-171 */
		unsigned c = (unsigned)(center * size);
		for (unsigned i = 0; i < size; ++i) {
			int xi = (int)i - (int)c;
			float xn = ((float)xi / ((float)size - 1.0f)) / fill;
			out[i] = expf(-10.0f * powf(xn, 2.0f));
		}
		return;
	}
	unsigned width = (unsigned)(fill * size);
	unsigned c = (unsigned)(center * size);
	int minPos = (int)(c - width / 2);
	int maxPos = minPos + (int)width;
	if (maxPos < minPos) { int t = minPos; minPos = maxPos; maxPos = t; }
	const float a0 = 0.215578948f, a1 = 0.416631580f, a2 = 0.277263158f, a3 = 0.083578947f, a4 = 0.006947368f;
	for (unsigned i = 0; i < size; ++i) {
		int xi = (int)i - minPos;
		float xn = (float)xi / ((float)width - 1.0f);
		if (xn > 0.999f || xn < 0.0001f) { out[i] = 0.0f; continue; }
		switch (type) {
		case ORC_WIN_HANNING:                                           /* :143-163 */
			out[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * (double)xn)));
			break;
		case ORC_WIN_SINE:                                              /* :173-193 */
			out[i] = (float)sin(M_PI * (double)xn);
			break;
		case ORC_WIN_LANCZOS: {                                         /* :195-221 */
			float arg = 2.0f * xn - 1.0f;
			if (arg == 0.0f) out[i] = 1.0f;
			else out[i] = (float)(sin(M_PI * (double)arg) / (M_PI * (double)arg));
			break;
		}
		case ORC_WIN_FLATTOP:                                           /* :224-253 */
			out[i] = a0 - a1 * (float)cos(2.0 * M_PI * (double)xn)
			            + a2 * (float)cos(4.0 * M_PI * (double)xn)
			            - a3 * (float)cos(6.0 * M_PI * (double)xn)
			            + a4 * (float)cos(8.0 * M_PI * (double)xn);
			break;
		case ORC_WIN_RECT:                                              /* :121-141 */
		default:
			out[i] = 1.0f;
			break;
		}
	}
}

/* CU:516-521 fillSinusoidalScanCorrectionCurve */
void orc_sinusoidal_curve(int A, float* out) {
	for (int k = 0; k < A; ++k) {
		float arg = (float)(1.0 - ((2.0 * (double)(float)k) / (double)(float)A));
		out[k] = (float)(((double)(float)A / M_PI) * acos((double)arg));
	}
}

/* ------------------------------------------------------------------ chain */

#define REAL double
#define SUFFIX 64
#include "oct_oracle_chain.inc"
#undef REAL
#undef SUFFIX

#define REAL float
#define SUFFIX 32
#define ORC_IS_FLOAT 1
#include "oct_oracle_chain.inc"
#undef ORC_IS_FLOAT
#undef REAL
#undef SUFFIX

int orc_process(const orc_params* p, const void* raw,
                const float* resample, const float* dispersion, const float* window,
                const float* ppBackground, double* meanLine, int determineFpn,
                float* out, double* complexOut, int precision) {
	if (precision == 32)
		return orc_process_impl32(p, raw, resample, dispersion, window, ppBackground, meanLine, determineFpn, out, complexOut);
	return orc_process_impl64(p, raw, resample, dispersion, window, ppBackground, meanLine, determineFpn, out, complexOut);
}

/* ------------------------------------------------------------------ post / display */

/* CU:743-755 */
void orc_postprocess_background(const float* processed, int halfN, int A, float* bg) {
	for (int z = 0; z < halfN; ++z) {
		float sum = 0;
		for (int i = 0; i < A; ++i) sum += processed[z + (size_t)i * halfN];
		bg[z] = sum / (float)A;
	}
}

/* CU:810-860 updateDisplayedBscanFrame (host wrapper CU:1273-1279 resets frameNr >= depth to 0) */
void orc_bscan_frame(const float* vol, int halfN, int A, int Btot, unsigned frameNr,
                     unsigned nFrames, int fn, float* disp) {
	const size_t F = (size_t)halfN * A;
	if (frameNr >= (unsigned)Btot) frameNr = 0;
	for (size_t i = 0; i < F; ++i) {
		if (nFrames > 1) {
			if (fn == 0) {
				int cnt = 0; float sum = 0;
				for (unsigned j = 0; j < nFrames; ++j) {
					unsigned f = frameNr + j;
					if (f < (unsigned)Btot) { sum += vol[f * F + (F - 1) - i]; cnt++; }
				}
				disp[i] = sum / (float)cnt;
			} else if (fn == 1) {
				float mx = 0;
				for (unsigned j = 0; j < nFrames; ++j) {
					unsigned f = frameNr + j;
					if (f < (unsigned)Btot) { float v = vol[f * F + (F - 1) - i]; if (mx < v) mx = v; }
				}
				disp[i] = mx;
			}
		} else {
			disp[i] = vol[frameNr * F + (F - 1) - i];
		}
	}
}

/* CU:862-912 updateDisplayedEnFaceViewFrame (host wrapper CU:1292-1303) */
void orc_enface_frame(const float* vol, int halfN, int A, int Btot, unsigned frameNr,
                      unsigned nFrames, int fn, float* disp) {
	const size_t E = (size_t)A * Btot;
	const unsigned W = (unsigned)halfN;
	if (frameNr >= W) frameNr = 0;
	for (size_t i = 0; i < E; ++i) {
		if (nFrames > 1) {
			if (fn == 0) {
				int cnt = 0; float sum = 0;
				for (unsigned j = 0; j < nFrames; ++j) {
					unsigned f = frameNr + j;
					if (f < W) { sum += vol[f + i * W]; cnt++; }
				}
				disp[(E - 1) - i] = sum / (float)cnt;
			} else if (fn == 1) {
				float mx = 0;
				for (unsigned j = 0; j < nFrames; ++j) {
					unsigned f = frameNr + j;
					if (f < W) { float v = vol[f + i * W]; if (mx < v) mx = v; }
				}
				disp[(E - 1) - i] = mx;
			}
		} else {
			disp[(E - 1) - i] = vol[frameNr + i * W];
		}
	}
}

static float saturatef_(float x) { if (!(x > 0.0f)) return 0.0f; if (x > 1.0f) return 1.0f; return x; }

/* CU:943-967 floatToOutput */
void orc_float_to_output(const float* in, size_t n, int bitDepth, void* out) {
	for (size_t i = 0; i < n; ++i) {
		float s = saturatef_(in[i]);
		if (bitDepth <= 8) ((uint8_t*)out)[i] = (uint8_t)((double)s * 255.0);
		else if (bitDepth <= 10) ((uint16_t*)out)[i] = (uint16_t)((double)s * 1023.0);
		else if (bitDepth <= 12) ((uint16_t*)out)[i] = (uint16_t)((double)s * 4095.0);
		else if (bitDepth <= 16) ((uint16_t*)out)[i] = (uint16_t)((double)s * 65535.0);
		else if (bitDepth <= 24) ((uint32_t*)out)[i] = (uint32_t)(s * 16777215.0f);
		else {
			float v = s * 4294967295.0f;   /* constant rounds to 2^32 in fp32; the GPU cvt saturates */
			((uint32_t*)out)[i] = v >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)v;
		}
	}
}
