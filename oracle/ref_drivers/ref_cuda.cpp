/*
 * ref_cuda.cpp -- headless C driver around the reference's UNMODIFIED cuda_code.cu
 * (compiled in place from /root/reference by oracle/Makefile with
 *  nvcc --use_fast_math -arch=sm_100, against oracle/shim for the Qt/GL includes).
 * It does what Processing::slot_start does (processing.cpp:136-229): fill the
 * OctAlgorithmParameters singleton, regenerate the host LUTs, initializeCuda(), then
 * octCudaPipeline() per buffer -- and reads d_processedBuffer (file-scope global, CU:98) back.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: the GPU-side parity target ("reference CUDA path")
 * and the same-box GPU baseline.  Runs only where a GPU exists (the B200 box).
 */
#include "kernels.h"
#include <chrono>
#include <cstdio>
#include <cstring>

extern float* d_processedBuffer;   /* cuda_code.cu:98 */
#ifndef REFCUDA_NO_MEANLINE_GLOBAL
extern cufftComplex* d_meanALine;  /* cuda_code.cu:84 */
#else
extern "C" int octb200_adapter_get_mean_line(float* reIm, int n);
extern "C" int octb200_adapter_sync();
#endif

struct refcuda_cfg {
	int samplesPerLine, ascansPerBscan, bscansPerBuffer, buffersPerVolume, bitDepth;
	int bitshift, bscanFlip, signalLogScaling, sinusoidalScanCorrection;
	float signalGrayscaleMin, signalGrayscaleMax, signalMultiplicator, signalAddend;
	int backgroundRemoval, rollingAverageWindowSize;
	int resampling, resamplingInterpolation;
	float c0, c1, c2, c3;
	int dispersionCompensation;
	float d0, d1, d2, d3;
	int windowing, windowType;
	float windowCenter, windowFillFactor;
	int fixedPatternNoiseRemoval, continuousFixedPatternNoiseDetermination, bscansForNoiseDetermination;
	int postProcessBackgroundRemoval;
	float postProcessBackgroundWeight, postProcessBackgroundOffset;
	int streamToHost, saveAs32bitFloat;
};

static size_t g_outFloatsPerBuffer = 0;

extern "C" int refcuda_configure(const refcuda_cfg* c) {
	OctAlgorithmParameters* p = OctAlgorithmParameters::getInstance();
	p->samplesPerLine = c->samplesPerLine; p->ascansPerBscan = c->ascansPerBscan;
	p->bscansPerBuffer = c->bscansPerBuffer; p->buffersPerVolume = c->buffersPerVolume; p->bitDepth = c->bitDepth;
	p->acquisitionParamsChanged = true;
	p->bitshift = c->bitshift; p->bscanFlip = c->bscanFlip; p->signalLogScaling = c->signalLogScaling;
	p->sinusoidalScanCorrection = c->sinusoidalScanCorrection;
	p->signalGrayscaleMin = c->signalGrayscaleMin; p->signalGrayscaleMax = c->signalGrayscaleMax;
	p->signalMultiplicator = c->signalMultiplicator; p->signalAddend = c->signalAddend;
	p->backgroundRemoval = c->backgroundRemoval; p->rollingAverageWindowSize = c->rollingAverageWindowSize;
	p->resampling = c->resampling;
	p->resamplingInterpolation = (OctAlgorithmParameters::INTERPOLATION)c->resamplingInterpolation;
	p->useCustomResampleCurve = false;
	p->c0 = c->c0; p->c1 = c->c1; p->c2 = c->c2; p->c3 = c->c3;
	p->dispersionCompensation = c->dispersionCompensation;
	p->d0 = c->d0; p->d1 = c->d1; p->d2 = c->d2; p->d3 = c->d3;
	p->windowing = c->windowing; p->window = (WindowFunction::WindowType)c->windowType;
	p->windowCenter = c->windowCenter; p->windowFillFactor = c->windowFillFactor;
	p->fixedPatternNoiseRemoval = c->fixedPatternNoiseRemoval;
	p->continuousFixedPatternNoiseDetermination = c->continuousFixedPatternNoiseDetermination;
	p->redetermineFixedPatternNoise = false;
	p->bscansForNoiseDetermination = c->bscansForNoiseDetermination;
	p->postProcessBackgroundRemoval = c->postProcessBackgroundRemoval;
	p->postProcessBackgroundWeight = c->postProcessBackgroundWeight;
	p->postProcessBackgroundOffset = c->postProcessBackgroundOffset;
	p->postProcessBackgroundRecordingRequested = false;
	p->bscanViewEnabled = false; p->enFaceViewEnabled = false; p->volumeViewEnabled = false; /* no GL here */
	p->streamToHost = c->streamToHost; p->streamingParamsChanged = false; p->streamingBuffersToSkip = 0;
	p->recParams.saveAs32bitFloat = c->saveAs32bitFloat;
	/* what Sidebar::slot_updateProcessingParams does (sidebar.cpp:463-473) */
	p->updateResampleCurve();
	p->updateDispersionCurve();
	p->updateWindowCurve();
	p->updatePostProcessingBackgroundCurve();
	p->acquisitionParamsChanged = false;
	g_outFloatsPerBuffer = (size_t)c->samplesPerLine / 2 * c->ascansPerBscan * c->bscansPerBuffer;
	return 0;
}

extern "C" int refcuda_get_curves(float* resample, float* dispersion, float* window) {
	OctAlgorithmParameters* p = OctAlgorithmParameters::getInstance();
	const size_t n = p->samplesPerLine;
	if (resample && p->resampleCurve) std::memcpy(resample, p->resampleCurve, n * sizeof(float));
	if (dispersion && p->dispersionCurve) std::memcpy(dispersion, p->dispersionCurve, n * sizeof(float));
	if (window && p->windowCurve) std::memcpy(window, p->windowCurve, n * sizeof(float));
	return 0;
}

extern "C" int refcuda_set_postprocess_background(const float* bg, int n) {
	OctAlgorithmParameters::getInstance()->loadPostProcessingBackground(const_cast<float*>(bg), n);
	return 0;
}
extern "C" void refcuda_request_background_recording() {
	OctAlgorithmParameters::getInstance()->postProcessBackgroundRecordingRequested = true;
}
extern "C" void refcuda_redetermine_fpn() {
	OctAlgorithmParameters::getInstance()->redetermineFixedPatternNoise = true;
}

extern "C" int refcuda_init(void* h1, void* h2) {
	return initializeCuda(h1, h2, OctAlgorithmParameters::getInstance()) ? 0 : -1;
}
extern "C" void refcuda_register_streaming(void* h1, void* h2, size_t bytes) { cuda_registerStreamingBuffers(h1, h2, bytes); }
extern "C" void refcuda_unregister_streaming() { cuda_unregisterStreamingBuffers(); }

/* h_in == NULL re-processes what is already in d_inputBuffer (CU:1400: the H2D copy is skipped) */
extern "C" void refcuda_process(void* h_in) { octCudaPipeline(h_in); }
extern "C" int refcuda_sync() { return (int)cudaDeviceSynchronize(); }

extern "C" int refcuda_copy_output(float* host, int bufferNrInVolume) {
	cudaDeviceSynchronize();
	return (int)cudaMemcpy(host, d_processedBuffer + g_outFloatsPerBuffer * bufferNrInVolume,
	                       g_outFloatsPerBuffer * sizeof(float), cudaMemcpyDeviceToHost);
}

/* wall-clock seconds for `iters` pipeline calls, device-synchronised on both sides */
extern "C" double refcuda_time(void* h_in_a, void* h_in_b, int iters, int warmup) {
	for (int i = 0; i < warmup; ++i) octCudaPipeline((i & 1) ? h_in_b : h_in_a);
	cudaDeviceSynchronize();
	auto t0 = std::chrono::steady_clock::now();
	for (int i = 0; i < iters; ++i) octCudaPipeline((i & 1) ? h_in_b : h_in_a);
	cudaDeviceSynchronize();
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

/* the fixed-pattern-noise mean line the reference determined (N complex values) */
extern "C" int refcuda_get_mean_line(float* reIm, int n) {
	cudaDeviceSynchronize();
#ifndef REFCUDA_NO_MEANLINE_GLOBAL
	return (int)cudaMemcpy(reIm, d_meanALine, sizeof(cufftComplex) * n, cudaMemcpyDeviceToHost);
#else
	return octb200_adapter_get_mean_line(reIm, n);
#endif
}

extern "C" void refcuda_cleanup() { cleanupCuda(); }
