/*
 * octb200.h -- C ABI of the B200-native OCT raw -> B-scan pipeline.
 *
 * This is the drop-in boundary for OCTproZ's GPU path.  It replaces the `extern "C"` set of
 * the reference's kernels.h (octproz_project/octproz/src/kernels.h:63-84, implemented in
 * cuda_code.cu) -- see the "replaces" note on every entry point -- but as a re-entrant,
 * handle-based, status-returning API with plain pointers and sizes: no Qt, no torch and no
 * CUDA types appear in the signatures (streams cross as void*).  The adapter that re-exports
 * the reference's own symbol names on top of this header is integration/octproz_kernels_adapter.cpp
 * (INTEGRATION.md).
 *
 * Data layout (identical to the reference):
 *   raw input   : [bscansPerBuffer][ascansPerBscan][samplesPerLine] containers, little endian,
 *                 container = u8 (bitDepth<=8), u16 (<=16) or u32 (else)   (cuda_code.cu:116-125)
 *   output      : float [buffersPerVolume*bscansPerBuffer][ascansPerBscan][samplesPerLine/2]
 *                 (cuda_code.cu:1118,1535: one slab per buffer of the volume)
 *
 * Threading: one host thread per pipeline handle at a time (as the reference: all calls come
 * from the processing thread, processing.cpp:176-218).  Handles are independent.
 * Errors: every call returns OCTB200_OK (0) or a negative code; octb200_last_error() has text.
 * Nothing here ever calls exit() (the reference's checkCudaErrors does, helper_cuda.h:582-595).
 */
#ifndef OCTB200_H
#define OCTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define OCTB200_API __declspec(dllexport)
#else
#define OCTB200_API __attribute__((visibility("default")))
#endif

#define OCTB200_VERSION 100

enum {
	OCTB200_OK = 0,
	OCTB200_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
	OCTB200_ERR_CUDA = -2,         /* CUDA runtime / cuFFT error, text in octb200_last_error */
	OCTB200_ERR_NOMEM = -3,        /* device allocation failed (initializeCuda returned false) */
	OCTB200_ERR_NOT_READY = -4     /* curves / buffers not set for an enabled stage */
};

/* octalgorithmparameters.h:55-59 */
enum { OCTB200_INTERP_LINEAR = 0, OCTB200_INTERP_CUBIC = 1, OCTB200_INTERP_LANCZOS = 2 };
/* windowfunction.h:41-48 */
enum { OCTB200_WIN_HANNING = 0, OCTB200_WIN_GAUSS = 1, OCTB200_WIN_SINE = 2,
       OCTB200_WIN_LANCZOS = 3, OCTB200_WIN_RECTANGULAR = 4, OCTB200_WIN_FLATTOP = 5 };
/* octalgorithmparameters.h:177-180 */
enum { OCTB200_DISPLAY_AVERAGING = 0, OCTB200_DISPLAY_MIP = 1 };

/* raw input packing.  CONTAINER = the reference's format (u8 / u16 / u32 containers, cuda_code.cu:116-125).  12P = an
   extension the reference does not have (docs/docs/faq.md: 12-bit data must be delivered in 16-bit containers): 12-bit samples
   packed little-endian, two samples per three bytes (GenICam PFNC "Mono12p": sample k of a line occupies bits [12k, 12k+12) of the
   line's bit string), bitDepth must be 12; a raw buffer then has samples*3/2 bytes -- a quarter less PCIe and HBM input traffic.
   The fused kernel unpacks in its slot conversion; Lanczos / rolling-mean / SPLIT / CUFFT chains unpack once into HBM first. */
enum { OCTB200_PACK_CONTAINER = 0, OCTB200_PACK_12P = 1 };

/* octb200_config.flags.  SEPARATE_CONVERSION: run floatToOutput (cuda_code.cu:943-967) as its own pass over the finished slab, as
   the reference does (cuda_code.cu:1366), instead of writing the converted u16 line from the fused kernel's epilogue (the default
   when the slab is final after that kernel; results are bit-identical, the flag exists for A/B measurement) */
enum { OCTB200_FLAG_SEPARATE_CONVERSION = 1,
       /* NO_DEPENDENT_LAUNCH: back-to-back buffers are normally launched with programmatic stream serialization (the next buffer's
          kernel prologue -- tensor-memory allocation, table fill, first line load -- overlaps the previous kernel's tail; results are
          identical).  The flag launches every kernel plainly (A/B measurement) */
       OCTB200_FLAG_NO_DEPENDENT_LAUNCH = 2 };

/* which kernels run the FFT stage */
enum {
	OCTB200_FFT_AUTO = 0,        /* FUSED wherever one of the fused kernels takes the geometry, else CUFFT */
	OCTB200_FFT_FUSED = 1,       /* ONE kernel: raw -> resample/window/phasor -> FFT -> FPN/log -> B-scan (container bytes in + 2 B out per
	                                sample).  samplesPerLine 1024 / 2048 in a u16 container: the transform runs in registers (k_fused.cuh);
	                                any other even samplesPerLine <= 8192 whose prime factors are <= 13 (e.g. the reference's default
	                                1664 = 2^7 * 13, 512, 1536, 4096) and u8 / u32 containers: in shared memory (k_generic.cu) */
	OCTB200_FFT_SPLIT = 2,       /* fused pre-FFT kernel -> float2 in HBM -> own FFT with fused epilogue (20 B/sample); 1024 / 2048 only */
	OCTB200_FFT_CUFFT = 3        /* fused pre-FFT kernel -> cufftExecC2C -> fused post kernel (32 B/sample); any N: the measured library
	                                baseline and the path for line lengths with larger prime factors */
};

/* AcquisitionParams (octproz_devkit/src/acquisitionparameter.h:31-37) + device placement */
typedef struct {
	uint32_t samplesPerLine;
	uint32_t ascansPerBscan;
	uint32_t bscansPerBuffer;
	uint32_t buffersPerVolume;
	uint32_t bitDepth;
	int32_t  device;            /* CUDA device ordinal; -1 = current device */
	int32_t  rawSlots;          /* device raw buffers for H2D / compute overlap; 0 = default (2) */
	int32_t  fftMode;           /* OCTB200_FFT_* */
	uint32_t bscanIndexBase;    /* multi-GPU shards: index (within the un-sharded buffer) of this shard's
	                               first B-scan, so "flip every even B-scan" (cuda_code.cu:795) keeps its parity */
	uint32_t inputPacking;      /* OCTB200_PACK_*: how the raw buffer stores its samples */
	uint32_t flags;             /* OCTB200_FLAG_* */
	uint32_t bscansInUnshardedBuffer; /* multi-GPU shards: B-scans per buffer of the un-sharded acquisition (0 = bscanIndexBase +
	                               bscansPerBuffer).  Only used by the B-scan flip: with an ODD number of B-scans per buffer the
	                               reference never flips the last one (cuda_bscanFlip covers samplesPerBuffer/4 elements,
	                               cuda_code.cu:794-805,1547), and only the shard that holds it can know that */
} octb200_config;

/* the [processing] block of OctAlgorithmParameters (octalgorithmparameters.h:108-166, 195-199) */
typedef struct {
	int32_t  bitshift;
	int32_t  bscanFlip;
	int32_t  signalLogScaling;
	int32_t  sinusoidalScanCorrection;
	float    signalGrayscaleMin;
	float    signalGrayscaleMax;
	float    signalMultiplicator;
	float    signalAddend;
	int32_t  backgroundRemoval;
	int32_t  rollingAverageWindowSize;
	int32_t  resampling;
	int32_t  resamplingInterpolation;      /* OCTB200_INTERP_* */
	int32_t  dispersionCompensation;
	int32_t  windowing;
	int32_t  fixedPatternNoiseRemoval;
	int32_t  continuousFixedPatternNoiseDetermination;
	int32_t  redetermineFixedPatternNoise;  /* edge trigger, consumed by the next process call (cuda_code.cu:1521-1524) */
	uint32_t bscansForNoiseDetermination;
	int32_t  postProcessBackgroundRemoval;
	int32_t  postProcessBackgroundRecordingRequested; /* edge trigger (cuda_code.cu:1558-1562) */
	float    postProcessBackgroundWeight;
	float    postProcessBackgroundOffset;
	int32_t  streamToHost;                  /* converted output -> registered host buffers + callback (cuda_code.cu:1601-1604) */
	uint32_t streamingBuffersToSkip;
	int32_t  streamFloatToHost;             /* recParams.saveAs32bitFloat path (cuda_code.cu:1596-1598) */
	uint32_t reserved[3];
} octb200_params;

typedef struct octb200_pipeline octb200_pipeline;

/* host callback: same shape as the Gpu2HostNotifier static callbacks (gpu2hostnotifier.h:47-49) */
typedef void (*octb200_host_callback)(void* hostBuffer);

/* ---------- lifecycle ---------- */
/* replaces initializeCuda (kernels.h:63, cuda_code.cu:1067-1162): allocates every device buffer, streams, events */
OCTB200_API int octb200_create(const octb200_config* cfg, octb200_pipeline** out);
/* replaces cleanupCuda / releaseBuffers / destroyStreamsAndEvents (kernels.h:65-67, cuda_code.cu:1164-1212) */
OCTB200_API int octb200_destroy(octb200_pipeline* p);
OCTB200_API const char* octb200_last_error(const octb200_pipeline* p);   /* p may be NULL: error of the calling thread's last failed create */
OCTB200_API int octb200_version(void);
OCTB200_API void octb200_default_params(octb200_params* out);           /* octalgorithmparameters.cpp:36-112 defaults */
OCTB200_API int octb200_effective_fft_mode(const octb200_pipeline* p);
/* which kernels OCTB200_FFT_AUTO runs the FFT stage of a geometry on -- host logic only, needs no GPU.  Returns one of OCTB200_PATH_*
   (negative: invalid geometry); for the shared-memory kernel `radices[0 .. *nPasses)` receives the plan of its Stockham passes
   (product = samplesPerLine; radices 13, 11, 7, 5, 3 first, then 8s, then 4 / 4·4 / 2).  radices (16 ints) and nPasses may be NULL.
   The reference plans cuFFT for any length (cuda_code.cu:1140); its default geometry is 1664 samples per line. */
enum { OCTB200_PATH_REGISTER_KERNEL = 1,                 /* 1024 / 2048: fused, transform in registers (u8 / u16 / u32 containers) */
       OCTB200_PATH_SHARED_MEMORY_KERNEL = 2,            /* fused, mixed-radix transform in shared memory */
       OCTB200_PATH_CUFFT_CHAIN = 3,                     /* pre kernel + cuFFT + post kernel: prime factors > 13 or > 8192 samples */
       OCTB200_PATH_CUFFT_CHAIN_SHARED_AVAILABLE = 4 };  /* short power-of-two lines: AUTO takes the cuFFT chain (as fast or faster on a
                                                            B200), an explicit OCTB200_FFT_FUSED gets the shared-memory kernel */
OCTB200_API int octb200_query_fft_path(uint32_t samplesPerLine, uint32_t bitDepth, int32_t* radices, int32_t* nPasses);

/* ---------- parameters and curves ---------- */
/* replaces the unsynchronised reads of the OctAlgorithmParameters singleton inside octCudaPipeline */
OCTB200_API int octb200_set_params(octb200_pipeline* p, const octb200_params* prm);
/* replace cuda_updateResampleCurve / cuda_updateDispersionCurve / cuda_updateWindowCurve /
   cuda_updatePostProcessBackground (cuda_code.cu:636-650,969-973): host fp32 LUTs of length n */
OCTB200_API int octb200_set_resample_curve(octb200_pipeline* p, const float* curve, int n);
OCTB200_API int octb200_set_dispersion_curve(octb200_pipeline* p, const float* phase, int n);
OCTB200_API int octb200_set_window_curve(octb200_pipeline* p, const float* window, int n);
OCTB200_API int octb200_set_postprocess_background(octb200_pipeline* p, const float* bg, int n);
OCTB200_API int octb200_get_postprocess_background(octb200_pipeline* p, float* bg, int n);
/* fixed-pattern-noise mean line (complex, n = samplesPerLine pairs).  Used by the multi-GPU host to
   broadcast rank 0's line; set marks the line as determined (cuda_code.cu:1523). */
OCTB200_API int octb200_get_fpn_mean_line(octb200_pipeline* p, float* reIm, int n);
OCTB200_API int octb200_set_fpn_mean_line(octb200_pipeline* p, const float* reIm, int n);
/* diagnostics of the last determination (getMinimumVarianceMean, cuda_code.cu:523-565, keeps per bin the mean of the segment with the
   smallest single-pass fp32 variance; at bins dominated by a constant term that minimum is decided by round-off): for the first
   `bins` (<= samplesPerLine/2) depth bins the NINE candidate segments, stats[(s * bins + z) * 4 + {0,1,2,3}] = { mean.re, mean.im,
   variance as computed (sumXX / L - |mean|^2 in fp32), mean power sumXX / L }; *segmentLength = L = height / 9.  Synchronises. */
OCTB200_API int octb200_get_fpn_segment_stats(octb200_pipeline* p, float* stats, int bins, int* segmentLength);

/* curve generators = OctAlgorithmParameters::update*Curve (octalgorithmparameters.cpp:141-249),
   Polynomial (polynomial.cpp:108-145), WindowFunction (windowfunction.cpp:121-253),
   fillSinusoidalScanCorrectionCurve (cuda_code.cu:516-521).  Pure host functions. */
OCTB200_API int octb200_make_resample_curve(int n, float c0, float c1, float c2, float c3, float* out);
OCTB200_API int octb200_make_dispersion_curve(int n, float d0, float d1, float d2, float d3, float* out);
OCTB200_API int octb200_make_window_curve(int type, float centerPosition, float fillFactor, int n, float* out);
OCTB200_API int octb200_make_sinusoidal_curve(int ascansPerBscan, float* out);

/* ---------- host buffers and callbacks ---------- */
/* replaces the cudaHostRegister of the plugin's two acquisition buffers (cuda_code.cu:1135-1136,1200-1207) */
OCTB200_API int octb200_register_host_buffers(octb200_pipeline* p, void* h1, void* h2);
OCTB200_API int octb200_unregister_host_buffers(octb200_pipeline* p);
/* replace cuda_register[Float]StreamingBuffers / cuda_unregister... (kernels.h:68-71) */
OCTB200_API int octb200_register_streaming_buffers(octb200_pipeline* p, void* h1, void* h2, size_t bytesPerBuffer);
OCTB200_API int octb200_unregister_streaming_buffers(octb200_pipeline* p);
OCTB200_API int octb200_register_float_streaming_buffers(octb200_pipeline* p, void* h1, void* h2, size_t bytesPerBuffer);
OCTB200_API int octb200_unregister_float_streaming_buffers(octb200_pipeline* p);
/* replace the hard-wired Gpu2HostNotifier::{dh2StreamingCallback, dh2FloatStreamingCallback, backgroundSignalCallback} */
OCTB200_API int octb200_set_callbacks(octb200_pipeline* p, octb200_host_callback streaming,
                                      octb200_host_callback floatStreaming, octb200_host_callback background);

/* ---------- the hot path ---------- */
/* replaces octCudaPipeline(void* h_inputSignal) (kernels.h:64, cuda_code.cu:1389-1605).
   h_raw: host buffer of one raw buffer.  The H2D copy runs on its own stream into one of `rawSlots`
   device slots; the call returns once the copy has finished (the producer may then refill h_raw, the
   contract of cuda_code.cu:1416-1419) while the kernels keep running asynchronously.
   h_raw == NULL re-processes the slot that was filled last (cuda_code.cu:1400). */
OCTB200_API int octb200_process_host(octb200_pipeline* p, const void* h_raw);
/* device-resident variant: d_raw is already in HBM (any 16-byte aligned device pointer) */
OCTB200_API int octb200_process_device(octb200_pipeline* p, const void* d_raw);
/* wait for everything issued so far (kernels, D2H streaming, callbacks) */
OCTB200_API int octb200_sync(octb200_pipeline* p);
OCTB200_API uint32_t octb200_current_buffer_nr(const octb200_pipeline* p);   /* params->currentBufferNr (cuda_code.cu:1602) */

/* ---------- results ---------- */
/* the processed volume in HBM (d_processedBuffer, cuda_code.cu:98); slab = buffer number in volume */
OCTB200_API float* octb200_output_device_ptr(octb200_pipeline* p, uint32_t bufferNrInVolume);
OCTB200_API int octb200_copy_output(octb200_pipeline* p, float* host, uint32_t bufferNrInVolume);
/* let the caller own the volume (e.g. a torch tensor): floats = buffersPerVolume * samplesPerBuffer/2; NULL restores the internal one */
OCTB200_API int octb200_bind_output(octb200_pipeline* p, void* d_volume);

/* replace changeDisplayedBscanFrame / changeDisplayedEnFaceFrame + updateDisplayed*Frame kernels
   (kernels.h:79-82, cuda_code.cu:810-912,1223-1308).  out is a DEVICE pointer (the mapped PBO in the
   Qt host): B-scan frame = N/2 * A floats, en-face frame = A * Btot floats. */
OCTB200_API int octb200_bscan_frame(octb200_pipeline* p, uint32_t frameNr, uint32_t displayFunctionFrames,
                                    int displayFunction, float* d_out);
OCTB200_API int octb200_enface_frame(octb200_pipeline* p, uint32_t frameNr, uint32_t displayFunctionFrames,
                                     int displayFunction, float* d_out);
/* replaces updateDisplayedVolume (cuda_code.cu:915-941): u8 voxels of one buffer into a linear device
   array laid out like the GL_R8 3-D texture: index = ((z * Btot) + y) * A + x, x = A-scan, y = B-scan in
   volume, z = (N/2-1) - depth */
OCTB200_API int octb200_volume_u8(octb200_pipeline* p, uint32_t bufferNrInVolume, uint8_t* d_out);
/* replaces floatToOutput (cuda_code.cu:943-967): saturate * (2^bits-1) into a device container array */
OCTB200_API int octb200_float_to_output(octb200_pipeline* p, uint32_t bufferNrInVolume, void* d_out);

/* ---------- multi-GPU: en-face frame of a B-scan-sharded volume, gathered over peer memory ----------
   The en-face view is the one product that needs every shard.  The reference has no multi-GPU path; what
   this replaces is "updateDisplayedEnFaceFrame per rank (cuda_code.cu:884-912) + one ncclAllGather": ONE kernel extracts
   each A-scan's en-face value and stores it directly into the frame window of every rank (P2P stores over NVLink), then
   publishes a per-rank sequence flag.  One process per GPU; the 64-byte handles are exchanged by the host
   (torch.distributed all_gather in octproz_b200/sharding.py).
     init    : allocate this rank's window for a volume of `globalLines` A-scans, this shard starting at line `lineOffset`
               (= first B-scan of the shard * ascansPerBscan); returns the window's IPC handle in handleOut[64]
     connect : handles = world * 64 bytes, rank-major; opens every peer window
     gather  : enqueue extraction + peer stores + flag publication on the compute stream, then the consumer side for the same
               sequence number: wait for ALL ranks' slabs, copy the assembled frame into this rank's private display frame and
               acknowledge to every producer -- in stream order, as the programmatic dependent of the producing kernel (its launch
               latency hides behind that kernel's tail).  The windows hold three frame buffers used round-robin by sequence number; a producer only overwrites a
               buffer after every rank has acknowledged the frame it held (flow control in the kernels' prologue), so a rank that
               runs ahead can never tear a frame a slower rank is still reading.  COLLECTIVE: every rank must issue the same
               sequence of gathers; a rank that stops gathering stalls its peers three gathers later -- for at most 10 s per launch:
               every device-side wait has a time-out that is counted (status) instead of hanging
     auto    : from now on EVERY octb200_process_* call also gathers frame (frameNr, frames, function) of the buffer it produced.
               When one depth frame is displayed and the slab is final after the fused kernel (single-slab volume, no sinusoidal
               correction, no background recording) the extraction and the peer stores happen inside that kernel's epilogue, spread
               over the whole kernel (a line group stores the values of up to 8 neighbouring lines with one coalesced store per
               rank; no end-of-kernel push, no grid-wide barrier) -- compute and collective in one launch; otherwise the stand-alone
               gather kernel is appended to the chain
     wait    : *dFrame = device pointer of this rank's display frame [globalLines] floats, reference order disp[(E-1)-i]; valid
               (stream ordered on the compute stream) until the next gather.  The consumer kernel was already enqueued by gather
     status  : sequence number of the latest gather and the number of time-outs so far (0 / 0 in a healthy run); synchronises
     close   : release (collective in spirit: call after a barrier, peers must have stopped gathering) */
#define OCTB200_IPC_HANDLE_BYTES 64
OCTB200_API int octb200_enface_gather_init(octb200_pipeline* p, int rank, int world, uint32_t globalLines, uint32_t lineOffset, void* handleOut);
OCTB200_API int octb200_enface_gather_connect(octb200_pipeline* p, const void* handles);
OCTB200_API int octb200_enface_gather(octb200_pipeline* p, uint32_t frameNr, uint32_t displayFunctionFrames, int displayFunction);
OCTB200_API int octb200_enface_gather_auto(octb200_pipeline* p, int enable, uint32_t frameNr, uint32_t displayFunctionFrames, int displayFunction);
OCTB200_API int octb200_enface_gather_wait(octb200_pipeline* p, float** dFrame);
OCTB200_API int octb200_enface_gather_status(octb200_pipeline* p, uint32_t* sequence, uint32_t* ackTimeouts, uint32_t* arrivalTimeouts);
OCTB200_API int octb200_enface_gather_close(octb200_pipeline* p);

/* ---------- dispersion-estimator sweep (octproz-dispersion-estimator-extension) ----------
   Replaces the loop of DispersionEstimationEngine::processDispersionMetric (src/dispersionestimationengine.cpp:78-90,118-158),
   which re-runs the CPU path (octprocessor/processor.tpp:241-321) once per trial coefficient, and
   AscanMetricCalculator::calculateMetric (src/ascanmetriccalculator.cpp:22-128): ALL trials are processed by ONE launch of the
   fused kernel (gridDim.y = trials, every trial reads the same center A-scans through its own window x phasor table) and one
   metric kernel.  Stage settings (resampling + interpolation, windowing, rolling background removal, bitshift) and the resample /
   window curves are the handle's; dispersion compensation is on with the trial coefficients; fixed-pattern-noise removal, flip,
   sinusoidal correction and background removal do not exist in the CPU path and are not applied.  The A-scans are scaled in the
   CPU path's units (processor.tpp:424-470), so thresholds mean what they mean in the reference.
   raw: `lines` A-scans (host or device memory, u16 containers).  coeffs: trials x {d0, d1, d2, d3} (octalgorithmparameters.cpp:
   dispersion polynomial).  metricsOut: host, `trials` floats.  ascansOut: NULL or host [trials][lines][samplesPerLine/2]. */
enum { OCTB200_METRIC_SUM_ABOVE_THRESHOLD = 0, OCTB200_METRIC_SAMPLES_ABOVE_THRESHOLD = 1, OCTB200_METRIC_PEAK_VALUE = 2,
       OCTB200_METRIC_MEAN_SOBEL = 3 };                /* ASCAN_SHARPNESS_METRIC, dispersionestimatorparameters.h:51-56 */
typedef struct {
	uint32_t lines;            /* numberOfCenterAscans */
	uint32_t trials;           /* numberOfDispersionSamples (<= 65535) */
	int32_t  metric;           /* OCTB200_METRIC_* */
	float    metricThreshold;
	int32_t  samplesToIgnore;  /* numberOfAscanSamplesToIgnore */
	int32_t  logScale;         /* !useLinearAscans (dispersionestimationengine.cpp:29-33) */
	float    logMin, logMax, logCoeff, logAddend;      /* processorcontroller.cpp:85-89 */
	uint32_t reserved[4];
} octb200_sweep_config;
OCTB200_API int octb200_dispersion_sweep(octb200_pipeline* p, const void* raw, const octb200_sweep_config* cfg, const float* coeffs,
                                         float* metricsOut, float* ascansOut);

/* ---------- timing helpers (CUDA events on the pipeline's compute stream) ---------- */
OCTB200_API void* octb200_compute_stream(octb200_pipeline* p);           /* cudaStream_t as void* */
OCTB200_API int octb200_event_record(octb200_pipeline* p, int slot);     /* slot 0..7 */
OCTB200_API int octb200_event_elapsed_ms(octb200_pipeline* p, int slotStart, int slotStop, float* ms);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
OCTB200_API uint64_t octb200_launch_count(const octb200_pipeline* p);
/* time ONE named stage in isolation with events: 0 = whole process_device, 1 = dominant kernel only */
OCTB200_API int octb200_time_kernel(octb200_pipeline* p, const void* d_raw, int iters, float* msPerIter);

#ifdef __cplusplus
}
#endif
#endif /* OCTB200_H */
