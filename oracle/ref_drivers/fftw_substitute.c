/*
 * fftw_substitute.c -- the five FFTW3 entry points the reference's CPU path calls
 * (octproz-dispersion-estimator-extension/src/octprocessor/processor.tpp:36-38,46-48,426),
 * implemented from scratch because libfftw3 is not in this image (only Windows DLLs and
 * fftw3.h are vendored by the reference; on Linux its .pro links the system -lfftw3).
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  Every report that times the reference CPU path says
 * "FFTW-API substitute".  Algorithm: out-of-place Stockham autosort, radix-4 passes plus one
 * radix-2 pass when log2(n) is odd, fp64, twiddles tabulated in the plan; naive O(n^2) DFT for
 * non power-of-two n.  Compiled against the reference's vendored fftw3.h for the types only.
 */
#include <fftw3.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct fftw_plan_s {
	int n, sign, pow2;
	fftw_complex* in;
	fftw_complex* out;
	double* tw;      /* exp(sign*2*pi*i*k/n), k<n, interleaved */
	double* work;    /* n complex */
};

void* fftw_malloc(size_t n) {
	void* p = NULL;
	if (posix_memalign(&p, 64, n ? n : 64) != 0) return NULL;
	return p;
}
void fftw_free(void* p) { free(p); }

fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out, int sign, unsigned flags) {
	(void)flags;
	struct fftw_plan_s* p = (struct fftw_plan_s*)calloc(1, sizeof(*p));
	p->n = n; p->sign = sign; p->in = in; p->out = out;
	p->pow2 = n > 0 && (n & (n - 1)) == 0;
	p->tw = (double*)fftw_malloc(sizeof(double) * 2 * (size_t)n);
	p->work = (double*)fftw_malloc(sizeof(double) * 2 * (size_t)n);
	for (int k = 0; k < n; ++k) {
		double a = (double)sign * 2.0 * M_PI * (double)k / (double)n;
		p->tw[2 * k] = cos(a); p->tw[2 * k + 1] = sin(a);
	}
	return p;
}

void fftw_destroy_plan(fftw_plan p) {
	if (!p) return;
	fftw_free(p->tw); fftw_free(p->work); free(p);
}

/* one Stockham radix-4 pass: n = 4*l*m ; x[(4j+q)... ] see loop */
static void pass4(int n, int l, int m, const double* restrict x, double* restrict y,
                  const double* restrict tw, int sign) {
	/* x viewed as [4][l][m] -> y as [l][4][m]; twiddle step n/(4l) */
	const int tstep = n / (4 * l);
	for (int j = 0; j < l; ++j) {
		const double w1r = tw[2 * (j * tstep)], w1i = tw[2 * (j * tstep) + 1];
		const double w2r = tw[2 * (2 * j * tstep)], w2i = tw[2 * (2 * j * tstep) + 1];
		const double w3r = tw[2 * (3 * j * tstep)], w3i = tw[2 * (3 * j * tstep) + 1];
		for (int k = 0; k < m; ++k) {
			const double* a = x + 2 * ((size_t)(0 * l + j) * m + k);
			const double* b = x + 2 * ((size_t)(1 * l + j) * m + k);
			const double* c = x + 2 * ((size_t)(2 * l + j) * m + k);
			const double* d = x + 2 * ((size_t)(3 * l + j) * m + k);
			double s0r = a[0] + c[0], s0i = a[1] + c[1];
			double s1r = a[0] - c[0], s1i = a[1] - c[1];
			double s2r = b[0] + d[0], s2i = b[1] + d[1];
			double s3r = b[0] - d[0], s3i = b[1] - d[1];
			/* multiply s3 by sign*i */
			double t3r = -(double)sign * s3i, t3i = (double)sign * s3r;
			double y0r = s0r + s2r, y0i = s0i + s2i;
			double y1r = s1r + t3r, y1i = s1i + t3i;
			double y2r = s0r - s2r, y2i = s0i - s2i;
			double y3r = s1r - t3r, y3i = s1i - t3i;
			double* o = y + 2 * ((size_t)(4 * j) * m + k);
			o[0] = y0r; o[1] = y0i;
			o[2 * m] = y1r * w1r - y1i * w1i;     o[2 * m + 1] = y1r * w1i + y1i * w1r;
			o[4 * m] = y2r * w2r - y2i * w2i;     o[4 * m + 1] = y2r * w2i + y2i * w2r;
			o[6 * m] = y3r * w3r - y3i * w3i;     o[6 * m + 1] = y3r * w3i + y3i * w3r;
		}
	}
}

static void pass2(int n, int l, int m, const double* restrict x, double* restrict y,
                  const double* restrict tw) {
	const int tstep = n / (2 * l);
	for (int j = 0; j < l; ++j) {
		const double wr = tw[2 * (j * tstep)], wi = tw[2 * (j * tstep) + 1];
		for (int k = 0; k < m; ++k) {
			const double* a = x + 2 * ((size_t)j * m + k);
			const double* b = x + 2 * ((size_t)(l + j) * m + k);
			double sr = a[0] + b[0], si = a[1] + b[1];
			double dr = a[0] - b[0], di = a[1] - b[1];
			double* o = y + 2 * ((size_t)(2 * j) * m + k);
			o[0] = sr; o[1] = si;
			o[2 * m] = dr * wr - di * wi; o[2 * m + 1] = dr * wi + di * wr;
		}
	}
}

void fftw_execute(const fftw_plan p) {
	const int n = p->n;
	double* in = (double*)p->in;
	double* out = (double*)p->out;
	if (!p->pow2) {
		for (int k = 0; k < n; ++k) {
			double sr = 0, si = 0;
			for (int j = 0; j < n; ++j) {
				int idx = (int)(((long long)j * k) % n);
				double wr = p->tw[2 * idx], wi = p->tw[2 * idx + 1];
				sr += in[2 * j] * wr - in[2 * j + 1] * wi;
				si += in[2 * j] * wi + in[2 * j + 1] * wr;
			}
			p->work[2 * k] = sr; p->work[2 * k + 1] = si;
		}
		memcpy(out, p->work, sizeof(double) * 2 * (size_t)n);
		return;
	}
	/* Stockham DIF: l = number of sub-transforms' twiddle groups, m = stride */
	const double* src = in;
	double* bufs[2] = { p->work, out };
	/* count passes so that the final result lands in `out` */
	int passes = 0; { int r = n; while (r > 1) { if (r % 4 == 0) r /= 4; else r /= 2; passes++; } }
	int which = (passes % 2 == 0) ? 0 : 1;  /* first destination */
	if (n == 1) { out[0] = in[0]; out[1] = in[1]; return; }
	int l = n, m = 1;
	while (l > 1) {
		double* dst = bufs[which];
		if (l % 4 == 0) { l /= 4; pass4(n, l, m, src, dst, p->tw, p->sign); m *= 4; }
		else            { l /= 2; pass2(n, l, m, src, dst, p->tw); m *= 2; }
		src = dst; which ^= 1;
	}
	if (src != out) memcpy(out, src, sizeof(double) * 2 * (size_t)n);
}
