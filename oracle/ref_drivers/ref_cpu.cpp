/*
 * ref_cpu.cpp -- C entry point around the reference's OWN CPU processing path,
 * OCTSignalProcessing::Processor<float>::processRawData
 * (octproz-dispersion-estimator-extension/src/octprocessor/processor.tpp:241-321), included
 * verbatim from /root/reference and linked against oracle/ref_drivers/fftw_substitute.c
 * (FFTW-API substitute -- libfftw3 is not in the image).  Configured exactly like
 * ProcessorController::processData (processorcontroller.cpp:103-148).
 *
 * TIMED CPU BASELINE ONLY (bench.py cpu_baseline / --impl reference).  Not a parity oracle:
 * its window / normalisation differ from the GPU path.
 *
 * threads == 1 : as shipped (single thread, one Processor for the whole buffer)
 * threads  > 1 : OpenMP over B-scans, one Processor per thread.
 */
#include "processor.h"
#include <cstring>
#include <cmath>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using Proc = OCTSignalProcessing::Processor<float>;

static void run_chunk(const void* raw, size_t totalSamples, int bitDepth, int N, int A,
                      int rollingWindow, int removeDC, int resample, int dispersion, int window, int logScale,
                      const float* c, const float* d, float coeff, float gmin, float gmax, float addend,
                      float* out) {
	Proc proc((size_t)N, (size_t)rollingWindow);   /* processorcontroller.cpp:116 */
	Proc::ProcessingOptions opt;
	opt.removeDC = removeDC; opt.resample = resample; opt.useCustomResamplingCurve = false;
	opt.compensateDispersion = dispersion; opt.applyWindow = window; opt.computeIFFT = true; opt.logScale = logScale;
	proc.setProcessingOptions(opt);
	proc.setDispersionCoefficients(std::vector<float>(d, d + 4));
	proc.setResamplingCoefficients(std::vector<float>(c, c + 4));
	proc.setLogScaleParameters(coeff, gmin, gmax, addend, false);
	std::vector<std::vector<std::vector<float>>> res;
	proc.processRawData(raw, totalSamples, bitDepth, (size_t)A, res);
	const size_t H = (size_t)N / 2;
	size_t l = 0;
	for (auto& frame : res)
		for (auto& line : frame) { std::memcpy(out + l * H, line.data(), sizeof(float) * H); ++l; }
}

extern "C" int refcpu_max_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

extern "C" int refcpu_process(const void* raw, int bitDepth, int N, int A, int B,
                              int rollingWindow, int removeDC, int resample, int dispersion, int window, int logScale,
                              const float* c, const float* d, float coeff, float gmin, float gmax, float addend,
                              int threads, float* out /* [B][A][N/2] */) {
	const size_t bytes = (size_t)std::ceil(bitDepth / 8.0);
	const size_t frameSamples = (size_t)N * A;
	if (threads <= 1) {
		run_chunk(raw, frameSamples * B, bitDepth, N, A, rollingWindow, removeDC, resample, dispersion, window, logScale,
		          c, d, coeff, gmin, gmax, addend, out);
		return 1;
	}
#ifdef _OPENMP
	#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
	for (int b = 0; b < B; ++b) {
		run_chunk((const char*)raw + (size_t)b * frameSamples * bytes, frameSamples, bitDepth, N, A,
		          rollingWindow, removeDC, resample, dispersion, window, logScale,
		          c, d, coeff, gmin, gmax, addend, out + (size_t)b * frameSamples / 2);
	}
	return threads;
}
