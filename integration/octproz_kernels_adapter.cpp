/*
 * octproz_kernels_adapter.cpp -- drop-in replacement for octproz/src/cuda_code.cu inside the OCTproZ build.
 *
 * It defines the extern "C" symbols of octproz/src/kernels.h:63-84 on top of the C ABI of liboctb200
 * (include/octb200.h), so processing.cpp, the DevKit and every plugin stay byte-for-byte unchanged:
 * in octproz.pro / pri/cuda.pri replace CUDA_SOURCES += src/cuda_code.cu by this file and link -loctb200.
 *
 * What it does per call is exactly the marshalling the reference does implicitly by reading the
 * OctAlgorithmParameters singleton inside octCudaPipeline (cuda_code.cu:1409-1604):
 *   - copy the processing block into the POD octb200_params,
 *   - honour the *Updated edge triggers (cuda_code.cu:1433-1445,1563-1566) by uploading the host LUTs,
 *   - write back the fields the reference mutates (cuda_code.cu:1435,1440,1444,1524,1561,1565,1602),
 *   - route the device->host notifications to Gpu2HostNotifier's static callbacks (gpu2hostnotifier.h:47-49).
 *
 * Build requirements: the reference's own headers (octalgorithmparameters.h, gpu2hostnotifier.h -> Qt) and, for the
 * three cuda_registerGlBuffer* entry points, CUDA-GL interop (define OCTB200_WITH_GL).  In this repository it is
 * compile- and run-checked against oracle/shim (no Qt/GL in the image): oracle/Makefile target _ref/libadapter_api.so,
 * tests/test_gpu_adapter.py.
 */
#include "octalgorithmparameters.h"
#include "gpu2hostnotifier.h"
#include "../include/octb200.h"

#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#ifdef OCTB200_WITH_GL
#include <cuda_gl_interop.h>
#else
typedef unsigned int GLuint;
#endif

namespace {
octb200_pipeline* g_p = nullptr;
OctAlgorithmParameters* g_params = nullptr;
bool g_initialized = false;
#ifdef OCTB200_WITH_GL
cudaGraphicsResource* g_glBscan = nullptr; cudaGraphicsResource* g_glEnFace = nullptr; cudaGraphicsResource* g_glVolume = nullptr;
uint8_t* g_volStage = nullptr; size_t g_volStageBytes = 0;      /* linear u8 image of the GL_R8 3-D texture: [N/2][Btot][A] */

/* updateVolumeDisplayBuffer (cuda_code.cu:1310-1355): the reference writes the voxels of the current buffer through a surface object,
   one byte per thread; here octb200_volume_u8 writes the slab into a linear staging image with a tiled transpose and one
   cudaMemcpy3DAsync moves that slab (x = all A-scans, y = the B-scans of this buffer, z = all depths) into the mapped array */
void update_volume_view(const OctAlgorithmParameters* q, cudaStream_t st) {
	if (!g_glVolume) return;
	const size_t A = q->ascansPerBscan, B = q->bscansPerBuffer, Btot = B * q->buffersPerVolume, H = q->samplesPerLine / 2;
	const size_t bytes = A * Btot * H;
	if (g_volStageBytes != bytes) {
		if (g_volStage) cudaFree(g_volStage);
		g_volStage = nullptr; g_volStageBytes = 0;
		if (cudaMalloc((void**)&g_volStage, bytes) != cudaSuccess) { printf("Cuda: volume view staging buffer allocation failed\n"); return; }
		cudaMemsetAsync(g_volStage, 0, bytes, st);
		g_volStageBytes = bytes;
	}
	const unsigned nr = octb200_current_buffer_nr(g_p);
	if (octb200_volume_u8(g_p, nr, g_volStage) != OCTB200_OK) { printf("Cuda error: %s\n", octb200_last_error(g_p)); return; }
	cudaArray_t arr = nullptr;
	if (cudaGraphicsMapResources(1, &g_glVolume, st) != cudaSuccess) return;                     /* cuda_map3dTexture, cuda_code.cu:1669-1680 */
	if (cudaGraphicsSubResourceGetMappedArray(&arr, g_glVolume, 0, 0) == cudaSuccess && arr) {
		cudaMemcpy3DParms c;
		std::memset(&c, 0, sizeof(c));
		c.srcPtr = make_cudaPitchedPtr(g_volStage, A, A, Btot);
		c.srcPos = make_cudaPos(0, nr * B, 0);
		c.dstArray = arr;
		c.dstPos = make_cudaPos(0, nr * B, 0);
		c.extent = make_cudaExtent(A, B, H);
		c.kind = cudaMemcpyDeviceToDevice;
		if (cudaMemcpy3DAsync(&c, st) != cudaSuccess) printf("Cuda: volume view copy failed: %s\n", cudaGetErrorString(cudaGetLastError()));
	}
	cudaGraphicsUnmapResources(1, &g_glVolume, st);
}
#endif

void on_background(void* hostLine) {
	/* cuda_copyPostProcessBackgroundToHost (cuda_code.cu:652-657): the host copy lives in params->postProcessBackground */
	if (g_params && g_params->postProcessBackground && hostLine)
		std::memcpy(g_params->postProcessBackground, hostLine, sizeof(float) * (g_params->samplesPerLine / 2));
	Gpu2HostNotifier::backgroundSignalCallback(g_params ? (void*)g_params->postProcessBackground : hostLine);
}

void marshal(const OctAlgorithmParameters* q, octb200_params* o) {
	octb200_default_params(o);
	o->bitshift = q->bitshift; o->bscanFlip = q->bscanFlip; o->signalLogScaling = q->signalLogScaling;
	o->sinusoidalScanCorrection = q->sinusoidalScanCorrection;
	o->signalGrayscaleMin = q->signalGrayscaleMin; o->signalGrayscaleMax = q->signalGrayscaleMax;
	o->signalMultiplicator = q->signalMultiplicator; o->signalAddend = q->signalAddend;
	o->backgroundRemoval = q->backgroundRemoval; o->rollingAverageWindowSize = q->rollingAverageWindowSize;
	o->resampling = q->resampling; o->resamplingInterpolation = (int)q->resamplingInterpolation;
	o->dispersionCompensation = q->dispersionCompensation; o->windowing = q->windowing;
	o->fixedPatternNoiseRemoval = q->fixedPatternNoiseRemoval;
	o->continuousFixedPatternNoiseDetermination = q->continuousFixedPatternNoiseDetermination;
	o->redetermineFixedPatternNoise = q->redetermineFixedPatternNoise;
	o->bscansForNoiseDetermination = q->bscansForNoiseDetermination;
	o->postProcessBackgroundRemoval = q->postProcessBackgroundRemoval;
	o->postProcessBackgroundRecordingRequested = q->postProcessBackgroundRecordingRequested;
	o->postProcessBackgroundWeight = q->postProcessBackgroundWeight; o->postProcessBackgroundOffset = q->postProcessBackgroundOffset;
	o->streamToHost = q->streamToHost && !q->streamingParamsChanged;      /* cuda_code.cu:1601 */
	o->streamingBuffersToSkip = q->streamingBuffersToSkip;
	o->streamFloatToHost = q->recParams.saveAs32bitFloat;                 /* cuda_code.cu:1596 */
}
}  // namespace

/* the reference keeps the processed volume in a file-scope global (cuda_code.cu:98); harnesses read it */
float* d_processedBuffer = nullptr;

extern "C" bool initializeCuda(void* h_buffer1, void* h_buffer2, OctAlgorithmParameters* parameters) {
	octb200_config cfg;
	std::memset(&cfg, 0, sizeof(cfg));
	cfg.samplesPerLine = parameters->samplesPerLine; cfg.ascansPerBscan = parameters->ascansPerBscan;
	cfg.bscansPerBuffer = parameters->bscansPerBuffer; cfg.buffersPerVolume = parameters->buffersPerVolume;
	cfg.bitDepth = parameters->bitDepth; cfg.device = -1; cfg.rawSlots = 2; cfg.fftMode = OCTB200_FFT_AUTO;
	if (octb200_create(&cfg, &g_p) != OCTB200_OK) {
		printf("octb200: %s\n", octb200_last_error(nullptr));             /* allocation failure -> false, as cuda_code.cu:1087-1129 */
		g_p = nullptr;
		return false;
	}
	g_params = parameters;
	if (h_buffer1 && octb200_register_host_buffers(g_p, h_buffer1, h_buffer2) != OCTB200_OK)     /* cuda_code.cu:1135-1136 */
		printf("octb200: %s\n", octb200_last_error(g_p));
	octb200_set_callbacks(g_p, Gpu2HostNotifier::dh2StreamingCallback, Gpu2HostNotifier::dh2FloatStreamingCallback, on_background);
	d_processedBuffer = octb200_output_device_ptr(g_p, 0);
	/* a fresh pipeline has no curves yet: force the first upload even if the GUI cleared the flags already */
	parameters->resamplingUpdated = parameters->resampling;
	parameters->dispersionUpdated = parameters->dispersionCompensation;
	parameters->windowUpdated = parameters->windowing;
	g_initialized = true;
	return true;
}

extern "C" void octCudaPipeline(void* h_inputSignal) {
	if (!g_initialized) { printf("Cuda: Device buffers are not initialized!\n"); return; }   /* cuda_code.cu:1391-1394 */
	OctAlgorithmParameters* q = g_params;
	const int N = (int)q->samplesPerLine;
	if (q->resampling && q->resamplingUpdated && q->resampleCurve) {                            /* cuda_code.cu:1433-1436 */
		octb200_set_resample_curve(g_p, q->resampleCurve, q->resampleCurveLength > 0 ? q->resampleCurveLength : N);
		q->resamplingUpdated = false;
	}
	if (q->dispersionCompensation && q->dispersionUpdated && q->dispersionCurve) {              /* cuda_code.cu:1437-1441 */
		octb200_set_dispersion_curve(g_p, q->dispersionCurve, N);
		q->dispersionUpdated = false;
	}
	if (q->windowing && q->windowUpdated && q->windowCurve) {                                    /* cuda_code.cu:1442-1445 */
		octb200_set_window_curve(g_p, q->windowCurve, N);
		q->windowUpdated = false;
	}
	if (q->postProcessBackgroundRemoval && q->postProcessBackgroundUpdated && q->postProcessBackground) {   /* cuda_code.cu:1563-1566 */
		octb200_set_postprocess_background(g_p, q->postProcessBackground, N / 2);
		q->postProcessBackgroundUpdated = false;
	}
	octb200_params prm;
	marshal(q, &prm);
	octb200_set_params(g_p, &prm);
	if (octb200_process_host(g_p, h_inputSignal) != OCTB200_OK)
		printf("Cuda error: %s\n", octb200_last_error(g_p));                                      /* cuda_code.cu:1590-1593: print only */
	/* fields the reference mutates */
	q->redetermineFixedPatternNoise = false;                                                      /* cuda_code.cu:1524 */
	q->postProcessBackgroundRecordingRequested = false;                                           /* cuda_code.cu:1561 */
	if (q->streamToHost && !q->streamingParamsChanged) q->currentBufferNr = octb200_current_buffer_nr(g_p);   /* cuda_code.cu:1602 */
#ifdef OCTB200_WITH_GL
	/* display buffers (cuda_code.cu:1571-1582): map the registered PBOs, let the library write into them, unmap */
	cudaStream_t st = (cudaStream_t)octb200_compute_stream(g_p);
	auto mapped = [&](cudaGraphicsResource* r) -> float* {
		if (!r) return nullptr;
		void* ptr = nullptr; size_t sz = 0;
		if (cudaGraphicsMapResources(1, &r, st) != cudaSuccess) return nullptr;
		cudaGraphicsResourceGetMappedPointer(&ptr, &sz, r);
		return (float*)ptr;
	};
	if (q->bscanViewEnabled) { if (float* d = mapped(g_glBscan)) { octb200_bscan_frame(g_p, q->frameNr, q->functionFramesBscan, q->displayFunctionBscan, d); cudaGraphicsUnmapResources(1, &g_glBscan, st); } }
	if (q->enFaceViewEnabled) { if (float* d = mapped(g_glEnFace)) { octb200_enface_frame(g_p, q->frameNrEnFaceView, q->functionFramesEnFaceView, q->displayFunctionEnFaceView, d); cudaGraphicsUnmapResources(1, &g_glEnFace, st); } }
	if (q->volumeViewEnabled) update_volume_view(q, st);                                          /* cuda_code.cu:1579-1582 */
#endif
}

extern "C" void releaseBuffers() {}
extern "C" void destroyStreamsAndEvents() {}
extern "C" void freeCudaMem(void** data) { if (data && *data) { cudaFree(*data); *data = nullptr; } }

extern "C" void cleanupCuda() {
	if (g_initialized) {                                                                          /* cuda_code.cu:1194-1212 */
		octb200_destroy(g_p);
		g_p = nullptr; g_initialized = false; d_processedBuffer = nullptr;
#ifdef OCTB200_WITH_GL
		if (g_volStage) { cudaFree(g_volStage); g_volStage = nullptr; g_volStageBytes = 0; }
#endif
	}
}

extern "C" void cuda_registerStreamingBuffers(void* h1, void* h2, size_t bytesPerBuffer) { if (g_p) octb200_register_streaming_buffers(g_p, h1, h2, bytesPerBuffer); }
extern "C" void cuda_unregisterStreamingBuffers() { if (g_p) octb200_unregister_streaming_buffers(g_p); }
extern "C" void cuda_registerFloatStreamingBuffers(void* h1, void* h2, size_t bytesPerBuffer) { if (g_p) octb200_register_float_streaming_buffers(g_p, h1, h2, bytesPerBuffer); }
extern "C" void cuda_unregisterFloatStreamingBuffers() { if (g_p) octb200_unregister_float_streaming_buffers(g_p); }

#ifdef OCTB200_WITH_GL
static bool reg_gl(cudaGraphicsResource** slot, GLuint buf) {
	if (*slot) { cudaGraphicsUnregisterResource(*slot); *slot = nullptr; }                       /* cuda_code.cu:1609-1615 */
	return cudaGraphicsGLRegisterBuffer(slot, buf, cudaGraphicsRegisterFlagsWriteDiscard) == cudaSuccess;
}
extern "C" bool cuda_registerGlBufferBscan(GLuint buf) { return reg_gl(&g_glBscan, buf); }
extern "C" bool cuda_registerGlBufferEnFaceView(GLuint buf) { return reg_gl(&g_glEnFace, buf); }
extern "C" bool cuda_registerGlBufferVolumeView(GLuint buf) {
	if (g_glVolume) { cudaGraphicsUnregisterResource(g_glVolume); g_glVolume = nullptr; }
	return cudaGraphicsGLRegisterImage(&g_glVolume, buf, 0x806F /* GL_TEXTURE_3D */, cudaGraphicsRegisterFlagsSurfaceLoadStore) == cudaSuccess;
}
#else
extern "C" bool cuda_registerGlBufferBscan(GLuint) { return false; }
extern "C" bool cuda_registerGlBufferEnFaceView(GLuint) { return false; }
extern "C" bool cuda_registerGlBufferVolumeView(GLuint) { return false; }
#endif

/* user-requested frame change while the stream is slow (cuda_code.cu:1223-1265): same kernels, user request path */
extern "C" void changeDisplayedBscanFrame(unsigned int frameNr, unsigned int displayFunctionFrames, int displayFunction) {
#ifdef OCTB200_WITH_GL
	if (!g_p || !g_glBscan) return;
	cudaStream_t st = (cudaStream_t)octb200_compute_stream(g_p);
	void* ptr = nullptr; size_t sz = 0;
	if (cudaGraphicsMapResources(1, &g_glBscan, st) != cudaSuccess) return;
	cudaGraphicsResourceGetMappedPointer(&ptr, &sz, g_glBscan);
	octb200_bscan_frame(g_p, frameNr, displayFunctionFrames, displayFunction, (float*)ptr);
	cudaGraphicsUnmapResources(1, &g_glBscan, st);
#else
	(void)frameNr; (void)displayFunctionFrames; (void)displayFunction;
#endif
}
extern "C" void changeDisplayedEnFaceFrame(unsigned int frameNr, unsigned int displayFunctionFrames, int displayFunction) {
#ifdef OCTB200_WITH_GL
	if (!g_p || !g_glEnFace) return;
	cudaStream_t st = (cudaStream_t)octb200_compute_stream(g_p);
	void* ptr = nullptr; size_t sz = 0;
	if (cudaGraphicsMapResources(1, &g_glEnFace, st) != cudaSuccess) return;
	cudaGraphicsResourceGetMappedPointer(&ptr, &sz, g_glEnFace);
	octb200_enface_frame(g_p, frameNr, displayFunctionFrames, displayFunction, (float*)ptr);
	cudaGraphicsUnmapResources(1, &g_glEnFace, st);
#else
	(void)frameNr; (void)displayFunctionFrames; (void)displayFunction;
#endif
}

/* harness hook (not part of kernels.h): FPN line of the adapter-owned pipeline */
extern "C" int octb200_adapter_get_mean_line(float* reIm, int n) { return g_p ? octb200_get_fpn_mean_line(g_p, reIm, n) : -1; }
extern "C" int octb200_adapter_get_fpn_segment_stats(float* stats, int bins, int* segmentLength) { return g_p ? octb200_get_fpn_segment_stats(g_p, stats, bins, segmentLength) : -1; }
extern "C" int octb200_adapter_sync() { return g_p ? octb200_sync(g_p) : -1; }
