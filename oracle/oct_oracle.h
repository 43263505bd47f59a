/*
 * oct_oracle.h -- CPU restatement of the OCTproZ raw->B-scan arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under octproz_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker.
 *
 * Parity status: the reference holds NO golden vectors or tests for this path
 * (SURVEY.md section 4 / 8c).  The LUT generators below are pinned against the
 * reference's own host code compiled verbatim (oracle/_ref/libref_luts.so, see
 * oracle/Makefile and tests/golden/make_golden.py).  The per-sample arithmetic is
 * pinned on the GPU box against the reference's unmodified cuda_code.cu compiled
 * for sm_100 (oracle/_ref/libref_cuda.so).  Where neither is available the
 * signal-chain parity is "unpinned" and says so in DESIGN.md.
 *
 * Every function cites the reference file:line it restates.  Paths are relative
 * to /root/reference/octproz_project/octproz/src/ (CU = cuda_code.cu).
 */
#ifndef OCT_ORACLE_H
#define OCT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* interpolation ids: octalgorithmparameters.h:55-59 */
enum { ORC_INTERP_LINEAR = 0, ORC_INTERP_CUBIC = 1, ORC_INTERP_LANCZOS = 2 };
/* window ids: windowfunction.h:41-48 */
enum { ORC_WIN_HANNING = 0, ORC_WIN_GAUSS = 1, ORC_WIN_SINE = 2, ORC_WIN_LANCZOS = 3,
       ORC_WIN_RECT = 4, ORC_WIN_FLATTOP = 5 };

typedef struct {
	int samplesPerLine;   /* N */
	int ascansPerBscan;   /* A */
	int bscansPerBuffer;  /* B */
	int bitDepth;         /* container: <=8 u8, <=16 u16, else u32 (CU:116-125) */
	int bitshift;
	int backgroundRemoval;
	int rollingAverageWindowSize;
	int resampling;
	int interpolation;
	int windowing;
	int dispersionCompensation;
	int fixedPatternNoiseRemoval;
	int bscansForNoiseDetermination;
	int signalLogScaling;
	float signalGrayscaleMin, signalGrayscaleMax, signalMultiplicator, signalAddend;
	int bscanFlip;
	int sinusoidalScanCorrection;
	int postProcessBackgroundRemoval;
	float postProcessBackgroundWeight, postProcessBackgroundOffset;
} orc_params;

/* ---- host LUT generators (fp32 results, like the reference) ---- */
/* polynomial.cpp:108-145 + octalgorithmparameters.cpp:141-167 */
void orc_resample_curve(int N, float c0, float c1, float c2, float c3, float* out);
/* octalgorithmparameters.cpp:206-222 (same polynomial form, no clamp) */
void orc_dispersion_curve(int N, float d0, float d1, float d2, float d3, float* out);
/* windowfunction.cpp:58-253 */
void orc_window_curve(int type, float center, float fill, int N, float* out);
/* CU:516-521 */
void orc_sinusoidal_curve(int A, float* out);

/* ---- signal chain ----
 * precision: 64 = all arithmetic in double (the yard-stick),
 *            32 = float arithmetic in the reference's operation order with libm
 *                 transcendentals (what an IEEE build of the reference computes).
 * raw          : [B][A][N] containers, little endian
 * resample/dispersion/window : fp32 LUTs of length N (may be NULL when the stage is off)
 * ppBackground : fp32 [N/2] or NULL
 * meanLine     : interleaved complex double [N] in/out.  If determineFpn != 0 the
 *                minimum-variance mean is (re)computed from this buffer first (CU:1521-1525).
 * out          : [B][A][N/2] float
 * complexOut   : optional [B][A][N] interleaved complex double (IFFT output before FPN), or NULL
 * returns 0 on success.
 */
int orc_process(const orc_params* p, const void* raw,
                const float* resample, const float* dispersion, const float* window,
                const float* ppBackground, double* meanLine, int determineFpn,
                float* out, double* complexOut, int precision);

/* CU:743-755: mean over the first A A-scans of a processed buffer, per depth bin */
void orc_postprocess_background(const float* processed, int halfN, int A, float* bg);

/* CU:810-860 / CU:862-912: display-frame extraction. fn: 0 averaging, 1 MIP */
void orc_bscan_frame(const float* vol, int halfN, int A, int Btot, unsigned frameNr,
                     unsigned nFrames, int fn, float* disp);
void orc_enface_frame(const float* vol, int halfN, int A, int Btot, unsigned frameNr,
                      unsigned nFrames, int fn, float* disp);
/* CU:943-967: saturate * (2^bits-1) -> container */
void orc_float_to_output(const float* in, size_t n, int bitDepth, void* out);

#ifdef __cplusplus
}
#endif
#endif
