/*
 * ref_luts.cpp -- C entry points around the reference's OWN host LUT code
 * (octalgorithmparameters.cpp, polynomial.cpp, windowfunction.cpp compiled verbatim from
 * /root/reference by oracle/Makefile against oracle/shim).  TEST INFRASTRUCTURE ONLY:
 * used here (container with /root/reference) to pin oracle/oct_oracle.c and the product's
 * curve generators, and to write tests/golden/*.npz.  Nothing of the reference is copied.
 */
#include "octalgorithmparameters.h"
#include <cstring>

extern "C" int ref_luts(int N,
                        float c0, float c1, float c2, float c3,
                        float d0, float d1, float d2, float d3,
                        int windowType, float center, float fill,
                        float* resample, float* dispersion, float* window) {
	OctAlgorithmParameters* p = OctAlgorithmParameters::getInstance();
	p->samplesPerLine = (unsigned)N;
	p->acquisitionParamsChanged = true;
	p->resampling = true; p->dispersionCompensation = true; p->windowing = true;
	p->useCustomResampleCurve = false;
	p->c0 = c0; p->c1 = c1; p->c2 = c2; p->c3 = c3;
	p->d0 = d0; p->d1 = d1; p->d2 = d2; p->d3 = d3;
	p->window = (WindowFunction::WindowType)windowType;
	p->windowCenter = center; p->windowFillFactor = fill;
	p->updateResampleCurve();
	p->updateDispersionCurve();
	p->updateWindowCurve();
	if (resample)   std::memcpy(resample,   p->resampleCurve,   sizeof(float) * N);
	if (dispersion) std::memcpy(dispersion, p->dispersionCurve, sizeof(float) * N);
	if (window)     std::memcpy(window,     p->windowCurve,     sizeof(float) * N);
	return 0;
}
