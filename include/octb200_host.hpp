/*
 * octb200_host.hpp -- the host side ABOVE the C ABI (include/octb200.h) in the reference's own language, without Qt.
 *
 * OCTproZ is a C++/Qt application; Qt is not available where this repository is built, so the classes a maintainer meets on the
 * path are mirrored here in plain C++17 with the reference's names, members and call order:
 *
 *   AcquisitionParams / AcquisitionBuffer   octproz_devkit/src/acquisitionparameter.h:31-37, acquisitionbuffer.h:43-68, .cpp:43-76
 *   AcquisitionSystem                        octproz_devkit/src/acquisitionsystem.h:58-73 (startAcquisition / stopAcquisition, public
 *                                            `buffer`, `params`, `acqusitionRunning` -- the reference's spelling; the Qt signals
 *                                            acquisitionStarted / acquisitionStopped are std::function members)
 *   VirtualOCTSystem                         octproz_plugins/octproz-virtual-oct-system/src/virtualoctsystem.cpp:59-224
 *   Gpu2HostNotifier                         octproz/src/gpu2hostnotifier.h:35-62 (process-wide receiver of the stream-to-host callbacks)
 *   OctPipeline                              the kernels.h entry points (kernels.h:63-84) as methods over an octb200 handle, incl. the
 *                                            streaming-buffer registration and the displayed-frame extraction into device memory
 *   Processing                               octproz/src/processing.cpp:124-229 (block the buffers, initializeCuda, poll the double
 *                                            buffer, octCudaPipeline, release the buffer, per-second statistics :194-207)
 *                                            and :231-266 slot_enableRecording (raw / processed recording sessions)
 *   Recorder / RecordingParams               octproz/src/recorder.cpp, octalgorithmparameters.h:84-98 (N buffers into one headerless file; session
 *                                            file naming, start with the first buffer of a volume, abort, meta file = settings INI copy)
 *   DispersionEstimationEngine               octproz-dispersion-estimator-extension/src/dispersionestimationengine.cpp:21-158 (the
 *                                            search; every sweep is one octb200_dispersion_sweep call instead of n CPU passes)
 *
 * Header only.  Processing is a template over the pipeline type so that the handshake can be exercised on a machine without a
 * GPU against a stand-in (tests/host/host_mirror_test.cpp); the product type is OctPipeline, which has no CPU fallback.
 * In a real OCTproZ build none of this is needed: integration/octproz_kernels_adapter.cpp re-exports the kernels.h names and the
 * Qt classes stay byte-for-byte what they are.
 */
#ifndef OCTB200_HOST_HPP
#define OCTB200_HOST_HPP

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "octb200.h"

namespace octb200 {
namespace host {

/* acquisitionparameter.h:31-37 */
struct AcquisitionParams {
	unsigned int samplesPerLine = 0;
	unsigned int ascansPerBscan = 0;
	unsigned int bscansPerBuffer = 0;
	unsigned int buffersPerVolume = 0;
	unsigned int bitDepth = 0;
};

/* acquisitionbuffer.{h,cpp}: bufferCnt 128-byte aligned host buffers + ready flags.  The reference's flags are plain bools polled by
 * two threads; here they are atomics with release / acquire so that the handshake is defined behaviour. */
class AcquisitionBuffer {
public:
	std::vector<void*> bufferArray;
	std::unique_ptr<std::atomic<bool>[]> bufferReadyArray;
	std::atomic<int> currIndex{-1};
	int bufferCnt = 0;
	size_t bytesPerBuffer = 0;

	AcquisitionBuffer() = default;
	AcquisitionBuffer(const AcquisitionBuffer&) = delete;
	AcquisitionBuffer& operator=(const AcquisitionBuffer&) = delete;
	~AcquisitionBuffer() { releaseMemory(); }

	bool allocateMemory(unsigned int count, size_t bytes) {            /* acquisitionbuffer.cpp:43-63 */
		releaseMemory();
		bufferReadyArray.reset(new std::atomic<bool>[count]);
		for (unsigned int i = 0; i < count; ++i) {
			void* p = nullptr;
			if (posix_memalign(&p, 128, bytes ? bytes : 128) != 0) { releaseMemory(); return false; }
			std::memset(p, 0, bytes);
			bufferArray.push_back(p);
			bufferReadyArray[i].store(false, std::memory_order_relaxed);
		}
		bufferCnt = (int)count;
		bytesPerBuffer = bytes;
		currIndex.store(-1);
		return true;
	}
	void releaseMemory() {                                              /* acquisitionbuffer.cpp:65-76 */
		for (void* p : bufferArray) std::free(p);
		bufferArray.clear();
		bufferReadyArray.reset();
		bufferCnt = 0;
		bytesPerBuffer = 0;
		currIndex.store(-1);
	}
	bool ready(int i) const { return bufferReadyArray[i].load(std::memory_order_acquire); }
	void setReady(int i, bool v) { bufferReadyArray[i].store(v, std::memory_order_release); }
};

/* acquisitionsystem.h:58-73 */
class AcquisitionSystem {
public:
	AcquisitionBuffer* buffer;
	AcquisitionParams params;
	std::atomic<bool> acqusitionRunning{false};                        /* sic */
	std::function<void(AcquisitionSystem*)> acquisitionStarted;        /* signal acquisitionStarted(AcquisitionSystem*) */
	std::function<void()> acquisitionStopped;                          /* signal acquisitionStopped() */

	AcquisitionSystem() : buffer(new AcquisitionBuffer()) {}
	virtual ~AcquisitionSystem() { delete buffer; }
	virtual void startAcquisition() = 0;
	virtual void stopAcquisition() { acqusitionRunning.store(false); }
};

/* virtualoctsystem.cpp: headerless little-endian raw file replayed into a two-slot buffer.  Settings = virtualoctsystemsettingsdialog.h:27-38. */
class VirtualOCTSystem : public AcquisitionSystem {
public:
	std::string filePath;
	int buffersFromFile = 2;
	unsigned int bscanOffset = 0;
	int waitTimeUs = 0;
	bool syncWithProcessing = true;
	std::atomic<long long> buffersDelivered{0};

	VirtualOCTSystem(const std::string& file, unsigned bitDepth, unsigned width, unsigned height, unsigned depth, unsigned buffersPerVolume = 1)
	    : filePath(file) {
		params.samplesPerLine = width; params.ascansPerBscan = height; params.bscansPerBuffer = depth;
		params.buffersPerVolume = buffersPerVolume; params.bitDepth = bitDepth;
	}

	size_t bytesOf(size_t samples) const { return samples * (size_t)std::ceil((double)params.bitDepth / 8.0); }   /* virtualoctsystem.cpp:124 */

	bool init() {                                                       /* virtualoctsystem.cpp:112-131 */
		FILE* f = std::fopen(filePath.c_str(), "rb");
		if (!f) return false;
		std::fclose(f);
		return buffer->allocateMemory(2, bytesOf((size_t)params.samplesPerLine * params.ascansPerBscan * params.bscansPerBuffer));
	}

	void startAcquisition() override {                                  /* virtualoctsystem.cpp:143-224 */
		if (!init()) { if (acquisitionStopped) acquisitionStopped(); return; }
		const size_t nBytes = buffer->bytesPerBuffer;
		const size_t offset = bytesOf((size_t)bscanOffset * params.samplesPerLine * params.ascansPerBscan);          /* :167 */
		FILE* f = std::fopen(filePath.c_str(), "rb");
		if (!f) { if (acquisitionStopped) acquisitionStopped(); return; }
		if (buffersFromFile > 2) {
			/* acqcuisitionSimulationLargeFile / acquisitionSimulationWithMultiFileBuffers (virtualoctsystem.cpp:107-113, 226-290): successive
			 * buffers of the file are streamed into the two acquisition buffers in turn, rewinding after buffersFromFile buffers */
			std::fseek(f, (long)offset, SEEK_SET);
			int readBuffers = 0, nxt = 0;
			acqusitionRunning.store(true);
			buffer->currIndex.store(1);
			if (acquisitionStarted) acquisitionStarted(this);
			while (acqusitionRunning.load()) {
				while (syncWithProcessing && buffer->ready(buffer->currIndex.load()) && acqusitionRunning.load()) std::this_thread::yield();
				if (!buffer->ready(nxt)) {
					const size_t got = std::fread(buffer->bufferArray[nxt], 1, nBytes, f); (void)got;
					if (++readBuffers >= buffersFromFile) { std::fseek(f, (long)offset, SEEK_SET); readBuffers = 0; }
					buffer->currIndex.store(nxt);
					buffer->setReady(nxt, true); buffersDelivered.fetch_add(1);
					nxt = (nxt + 1) % 2;
				}
				if (waitTimeUs > 0) std::this_thread::sleep_for(std::chrono::microseconds(waitTimeUs));
			}
			std::fclose(f);
			if (acquisitionStopped) acquisitionStopped();
			return;
		}
		auto readAt = [&](size_t off, void* dst) {
			if (std::fseek(f, (long)off, SEEK_SET) == 0) { const size_t got = std::fread(dst, 1, nBytes, f); (void)got; }
		};
		readAt(offset, buffer->bufferArray[0]);
		readAt(offset + (buffersFromFile == 2 ? nBytes : 0), buffer->bufferArray[1]);                                  /* :175-179 */
		std::fclose(f);
		acqusitionRunning.store(true);
		buffer->currIndex.store(1);
		if (acquisitionStarted) acquisitionStarted(this);
		while (acqusitionRunning.load()) {                              /* :196-223 */
			while (syncWithProcessing && buffer->ready(buffer->currIndex.load()) && acqusitionRunning.load()) std::this_thread::yield();
			const int nxt = (buffer->currIndex.load() + 1) % 2;
			buffer->currIndex.store(nxt);
			if (!buffer->ready(nxt)) { buffer->setReady(nxt, true); buffersDelivered.fetch_add(1); }
			if (waitTimeUs > 0) std::this_thread::sleep_for(std::chrono::microseconds(waitTimeUs));
		}
		if (acquisitionStopped) acquisitionStopped();
	}
};

/* The processing block of OctAlgorithmParameters as the C ABI carries it, plus the three host LUTs (octalgorithmparameters.h:108-166). */
struct OctAlgorithmParameters {
	octb200_params p;
	std::vector<float> resampleCurve, dispersionCurve, windowCurve, postProcessBackground;
	float c[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
	int window = OCTB200_WIN_HANNING;
	float windowCenter = 0.5f, windowFillFactor = 0.95f;

	OctAlgorithmParameters() { octb200_default_params(&p); }

	/* updateResampleCurve / updateDispersionCurve / updateWindowCurve (octalgorithmparameters.cpp:141-249) */
	bool updateCurves(unsigned int samplesPerLine) {
		const int n = (int)samplesPerLine;
		resampleCurve.resize(n); dispersionCurve.resize(n); windowCurve.resize(n);
		return octb200_make_resample_curve(n, c[0], c[1], c[2], c[3], resampleCurve.data()) == OCTB200_OK &&
		       octb200_make_dispersion_curve(n, d[0], d[1], d[2], d[3], dispersionCurve.data()) == OCTB200_OK &&
		       octb200_make_window_curve(window, windowCenter, windowFillFactor, n, windowCurve.data()) == OCTB200_OK;
	}

	/* Settings file of the reference (QSettings INI): the [processing] keys of octproz/src/sidebar.h:58-94, the
	 * [Virtual%20OCT%20System] keys of virtualoctsystemsettingsdialog.h:27-38 and [streaming].  Missing keys keep
	 * the defaults of the reference's settings code.  acq / vos receive the acquisition geometry and the replay settings when given. */
	struct VirtualOctSettings { std::string filePath; int buffersFromFile = 2; unsigned int bscanOffset = 0; int waitTimeUs = 0; bool syncWithProcessing = true; };
	static bool fromIni(const std::string& path, OctAlgorithmParameters* q, AcquisitionParams* acq = nullptr, VirtualOctSettings* vos = nullptr) {
		FILE* f = std::fopen(path.c_str(), "r");
		if (!f) return false;
		std::string section;
		char line[4096];
		auto trim = [](std::string t) {
			const char* ws = " \t\r\n";
			const size_t a = t.find_first_not_of(ws);
			if (a == std::string::npos) return std::string();
			return t.substr(a, t.find_last_not_of(ws) - a + 1);
		};
		auto truth = [](const std::string& v) { return v == "true" || v == "1" || v == "True"; };
		octb200_params& p = q->p;
		while (std::fgets(line, sizeof(line), f)) {
			const std::string t = trim(line);
			if (t.empty() || t[0] == ';' || t[0] == '#') continue;
			if (t[0] == '[') { section = t.substr(1, t.find(']') - 1); continue; }
			const size_t eq = t.find('=');
			if (eq == std::string::npos) continue;
			const std::string k = trim(t.substr(0, eq)), v = trim(t.substr(eq + 1));
			const double num = std::atof(v.c_str());
			if (section == "processing") {
				if (k == "bitshift") p.bitshift = truth(v);
				else if (k == "flip_bscans") p.bscanFlip = truth(v);
				else if (k == "log") p.signalLogScaling = truth(v);
				else if (k == "sinusoidal_scan_correction") p.sinusoidalScanCorrection = truth(v);
				else if (k == "max") p.signalGrayscaleMax = (float)num;
				else if (k == "min") p.signalGrayscaleMin = (float)num;
				else if (k == "coeff") p.signalMultiplicator = (float)num;
				else if (k == "addend") p.signalAddend = (float)num;
				else if (k == "background_removal") p.backgroundRemoval = truth(v);
				else if (k == "background_removal_window_size") p.rollingAverageWindowSize = (int)num;
				else if (k == "resampling") p.resampling = truth(v);
				else if (k == "resampling_interpolation") p.resamplingInterpolation = (int)num;
				else if (k.rfind("resampling_c", 0) == 0 && k.size() == 13) q->c[k[12] - '0'] = (float)num;
				else if (k == "dispersion_compensation") p.dispersionCompensation = truth(v);
				else if (k.rfind("dispersion_compensation_d", 0) == 0 && k.size() == 26) q->d[k[25] - '0'] = (float)num;
				else if (k == "windowing") p.windowing = truth(v);
				else if (k == "window_type") q->window = (int)num;
				else if (k == "window_fill_factor") q->windowFillFactor = (float)num;
				else if (k == "window_center_position") q->windowCenter = (float)num;
				else if (k == "fixed_pattern_removal") p.fixedPatternNoiseRemoval = truth(v);
				else if (k == "fixed_pattern_removal_continuously") p.continuousFixedPatternNoiseDetermination = truth(v);
				else if (k == "fixed_pattern_removal_bscans") p.bscansForNoiseDetermination = (uint32_t)num;
				else if (k == "post_processing_background_removal") p.postProcessBackgroundRemoval = truth(v);
				else if (k == "post_processing_background_removal_weight") p.postProcessBackgroundWeight = (float)num;
				else if (k == "post_processing_background_removal_offset") p.postProcessBackgroundOffset = (float)num;
			} else if (section == "Virtual%20OCT%20System" || section == "Virtual OCT System") {
				if (acq) {
					if (k == "bit_depth") acq->bitDepth = (unsigned)num;
					else if (k == "width") acq->samplesPerLine = (unsigned)num;
					else if (k == "height") acq->ascansPerBscan = (unsigned)num;
					else if (k == "depth") acq->bscansPerBuffer = (unsigned)num;
					else if (k == "buffers_per_volume") acq->buffersPerVolume = (unsigned)num;
				}
				if (vos) {
					if (k == "file_path") vos->filePath = v;
					else if (k == "buffers_from_file") vos->buffersFromFile = (int)num;
					else if (k == "bscan_offset") vos->bscanOffset = (unsigned)num;
					else if (k == "wait_time") vos->waitTimeUs = (int)num;
					else if (k == "sync_with_processing") vos->syncWithProcessing = truth(v);
				}
			} else if (section == "streaming") {
				if (k == "streaming_enabled") p.streamToHost = truth(v);
				else if (k == "streaming_skip") p.streamingBuffersToSkip = (uint32_t)num;
			}
		}
		std::fclose(f);
		return true;
	}

	/* the published benchmark settings (performance/v180/.../20250504_octproz_settings.ini:17-67) */
	static OctAlgorithmParameters benchmark(unsigned int samplesPerLine) {
		OctAlgorithmParameters q;
		q.p.signalLogScaling = 1; q.p.signalGrayscaleMin = -30.0f; q.p.signalGrayscaleMax = 100.0f; q.p.signalMultiplicator = 1.0f; q.p.signalAddend = 0.0f;
		q.p.resampling = 1; q.p.resamplingInterpolation = OCTB200_INTERP_CUBIC;
		const float s = (float)((double)samplesPerLine / 1024.0);
		q.c[0] = 0.535239f; q.c[1] = (float)(871.817574 * s); q.c[2] = (float)(-170.633784 * s); q.c[3] = (float)(97.249716 * s);
		q.p.dispersionCompensation = 1; q.d[0] = 0.0f; q.d[1] = 97.0f; q.d[2] = -96.625f; q.d[3] = -0.375f;
		q.p.windowing = 1; q.window = OCTB200_WIN_HANNING; q.windowFillFactor = 0.95f; q.windowCenter = 0.5f;
		q.p.fixedPatternNoiseRemoval = 1; q.p.bscansForNoiseDetermination = 1;
		q.updateCurves(samplesPerLine);
		return q;
	}
};

/* Gpu2HostNotifier (octproz/src/gpu2hostnotifier.h:35-62): the C callbacks carry no user pointer, so -- as in the reference -- one
 * process-wide object receives them and forwards to whoever connected (Qt signals there, std::function here).  The callbacks arrive on a
 * CUDA host-function thread; connect() / disconnect() are serialised with them. */
class Gpu2HostNotifier {
public:
	using Slot = std::function<void(void*)>;
	static Gpu2HostNotifier& getInstance() { static Gpu2HostNotifier n; return n; }
	void connectProcessed(Slot s) { std::lock_guard<std::recursive_mutex> g(m_); newGpuDataAvailable_ = std::move(s); }
	void connectFloat(Slot s) { std::lock_guard<std::recursive_mutex> g(m_); newGpuFloatDataAvailable_ = std::move(s); }
	void connectBackground(Slot s) { std::lock_guard<std::recursive_mutex> g(m_); backgroundRecorded_ = std::move(s); }
	void disconnectAll() { std::lock_guard<std::recursive_mutex> g(m_); newGpuDataAvailable_ = nullptr; newGpuFloatDataAvailable_ = nullptr; backgroundRecorded_ = nullptr; }
	/* the three functions handed to octb200_set_callbacks (gpu2hostnotifier.h:47-49) */
	static void dh2StreamingCallback(void* buffer) { getInstance().emit(&Gpu2HostNotifier::newGpuDataAvailable_, buffer); }
	static void dh2FloatStreamingCallback(void* buffer) { getInstance().emit(&Gpu2HostNotifier::newGpuFloatDataAvailable_, buffer); }
	static void backgroundSignalCallback(void* buffer) { getInstance().emit(&Gpu2HostNotifier::backgroundRecorded_, buffer); }

private:
	Gpu2HostNotifier() = default;
	void emit(Slot Gpu2HostNotifier::*which, void* buffer) {
		std::lock_guard<std::recursive_mutex> g(m_);
		if (this->*which) (this->*which)(buffer);
	}
	std::recursive_mutex m_;          /* a slot may (dis)connect from inside a callback */
	Slot newGpuDataAvailable_, newGpuFloatDataAvailable_, backgroundRecorded_;
};

/* kernels.h:63-84 over one octb200 handle.  Every failure throws std::runtime_error with octb200_last_error(): loud, no fallback. */
class OctPipeline {
public:
	explicit OctPipeline(int fftMode = OCTB200_FFT_AUTO, int device = -1) : fftMode_(fftMode), device_(device) {}
	OctPipeline(const OctPipeline&) = delete;
	OctPipeline& operator=(const OctPipeline&) = delete;
	~OctPipeline() { cleanupCuda(); }

	/* initializeCuda(void* h_buffer1, void* h_buffer2, OctAlgorithmParameters*) -- kernels.h:63, cuda_code.cu:1067-1162 */
	bool initializeCuda(void* hBuffer1, void* hBuffer2, const AcquisitionParams& acq, OctAlgorithmParameters* params) {
		cleanupCuda();
		octb200_config cfg;
		std::memset(&cfg, 0, sizeof(cfg));
		cfg.samplesPerLine = acq.samplesPerLine; cfg.ascansPerBscan = acq.ascansPerBscan; cfg.bscansPerBuffer = acq.bscansPerBuffer;
		cfg.buffersPerVolume = acq.buffersPerVolume ? acq.buffersPerVolume : 1; cfg.bitDepth = acq.bitDepth;
		cfg.device = device_; cfg.fftMode = fftMode_;
		if (octb200_create(&cfg, &h_) != OCTB200_OK) { lastError_ = octb200_last_error(nullptr); h_ = nullptr; return false; }
		acq_ = acq; params_ = params;
		if (hBuffer1 && hBuffer2) check(octb200_register_host_buffers(h_, hBuffer1, hBuffer2), "register_host_buffers");
		pushParams(true);
		return true;
	}
	/* octCudaPipeline(void* h_inputSignal) -- kernels.h:64, cuda_code.cu:1389-1605 */
	void octCudaPipeline(void* hInputSignal) {
		pushParams(false);
		check(octb200_process_host(h_, hInputSignal), "process_host");
	}
	/* cleanupCuda() -- kernels.h:65, cuda_code.cu:1164-1212 */
	void cleanupCuda() {
		if (h_) { octb200_destroy(h_); h_ = nullptr; }
	}
	void sync() { check(octb200_sync(h_), "sync"); }
	/* kernels.h:71-76: the two host buffers the converted (or float) output is streamed into, delivered through Gpu2HostNotifier */
	void cuda_registerStreamingBuffers(void* h1, void* h2, size_t bytesPerBuffer) {
		check(octb200_register_streaming_buffers(h_, h1, h2, bytesPerBuffer), "register_streaming_buffers"); connectNotifier();
	}
	void cuda_unregisterStreamingBuffers() { check(octb200_unregister_streaming_buffers(h_), "unregister_streaming_buffers"); }
	void cuda_registerFloatStreamingBuffers(void* h1, void* h2, size_t bytesPerBuffer) {
		check(octb200_register_float_streaming_buffers(h_, h1, h2, bytesPerBuffer), "register_float_streaming_buffers"); connectNotifier();
	}
	void cuda_unregisterFloatStreamingBuffers() { check(octb200_unregister_float_streaming_buffers(h_), "unregister_float_streaming_buffers"); }
	/* kernels.h:66-67 (without the GL mapping: the frame goes to caller-provided device memory) */
	void changeDisplayedBscanFrame(unsigned frameNr, unsigned displayFunctionFrames, int displayFunction, float* dOut) {
		check(octb200_bscan_frame(h_, frameNr, displayFunctionFrames, displayFunction, dOut), "bscan_frame");
	}
	void changeDisplayedEnFaceFrame(unsigned frameNr, unsigned displayFunctionFrames, int displayFunction, float* dOut) {
		check(octb200_enface_frame(h_, frameNr, displayFunctionFrames, displayFunction, dOut), "enface_frame");
	}
	unsigned currentBufferNr() const { return h_ ? octb200_current_buffer_nr(h_) : 0u; }                /* params->currentBufferNr, cuda_code.cu:1602 */
	void copyOutput(float* host, unsigned bufferNrInVolume = 0) { check(octb200_copy_output(h_, host, bufferNrInVolume), "copy_output"); }
	unsigned long long launchCount() const { return h_ ? (unsigned long long)octb200_launch_count(h_) : 0ULL; }
	octb200_pipeline* handle() { return h_; }
	const std::string& lastError() const { return lastError_; }

	/* the *Updated edge triggers of the reference (cuda_code.cu:1433-1445): callers flip these after changing a curve */
	bool resamplingUpdated = false, dispersionUpdated = false, windowUpdated = false, postProcessBackgroundUpdated = false;

private:
	void check(int rc, const char* what) {
		if (rc != OCTB200_OK) {
			lastError_ = std::string(what) + " failed (" + std::to_string(rc) + "): " + (octb200_last_error(h_) ? octb200_last_error(h_) : "");
			throw std::runtime_error(lastError_);
		}
	}
	void connectNotifier() {
		check(octb200_set_callbacks(h_, &Gpu2HostNotifier::dh2StreamingCallback, &Gpu2HostNotifier::dh2FloatStreamingCallback,
		                            &Gpu2HostNotifier::backgroundSignalCallback), "set_callbacks");
	}
	void pushParams(bool forceCurves) {
		OctAlgorithmParameters& q = *params_;
		check(octb200_set_params(h_, &q.p), "set_params");
		const int n = (int)acq_.samplesPerLine;
		if (q.p.resampling && (forceCurves || resamplingUpdated)) { check(octb200_set_resample_curve(h_, q.resampleCurve.data(), n), "set_resample_curve"); resamplingUpdated = false; }
		if (q.p.dispersionCompensation && (forceCurves || dispersionUpdated)) { check(octb200_set_dispersion_curve(h_, q.dispersionCurve.data(), n), "set_dispersion_curve"); dispersionUpdated = false; }
		if (q.p.windowing && (forceCurves || windowUpdated)) { check(octb200_set_window_curve(h_, q.windowCurve.data(), n), "set_window_curve"); windowUpdated = false; }
		if (q.p.postProcessBackgroundRemoval && postProcessBackgroundUpdated && !q.postProcessBackground.empty()) {
			check(octb200_set_postprocess_background(h_, q.postProcessBackground.data(), (int)q.postProcessBackground.size()), "set_postprocess_background");
			postProcessBackgroundUpdated = false;
		}
		/* edge triggers are consumed by the pipeline (cuda_code.cu:1524,1561) */
		q.p.redetermineFixedPatternNoise = 0;
		q.p.postProcessBackgroundRecordingRequested = 0;
	}

	int fftMode_, device_;
	octb200_pipeline* h_ = nullptr;
	AcquisitionParams acq_;
	OctAlgorithmParameters* params_ = nullptr;
	std::string lastError_;
};

/* OctAlgorithmParameters::RecordingParams (octproz/src/octalgorithmparameters.h:84-98) without the GUI-only screenshot switch */
struct RecordingParams {
	std::string timestamp, fileName, savePath;
	size_t bufferSizeInBytes = 0;
	unsigned int buffersToRecord = 0;
	bool startWithFirstBuffer = false, recordRaw = false, recordProcessed = false, saveMetaData = false, saveAs32bitFloat = false, stopAfterRecord = false;

	/* <savePath>/<timestamp>[_<fileName>]: shared by every file of one recording session (recorder.cpp:77-82, octprozapp.cpp:296) */
	std::string sessionPrefix() const { return savePath + "/" + timestamp + (fileName.empty() ? std::string() : "_" + fileName); }
	/* what Processing::slot_enableRecording hands to the processed-data recorder (processing.cpp:243-249) */
	RecordingParams forProcessedData(const AcquisitionParams& a) const {
		RecordingParams r = *this;
		r.bufferSizeInBytes = saveAs32bitFloat ? (size_t)(a.samplesPerLine / 2) * a.ascansPerBscan * a.bscansPerBuffer * sizeof(float) : bufferSizeInBytes / 2;
		return r;
	}
	/* the meta file of a recording is a copy of the settings INI (octprozapp.cpp:294-298); returns its path, empty if not written */
	std::string saveMeta(const std::string& settingsFile) const {
		if (!saveMetaData) return {};
		const std::string dst = sessionPrefix() + "_meta.txt";
		FILE* in = std::fopen(settingsFile.c_str(), "rb");
		if (!in) return {};
		FILE* out = std::fopen(dst.c_str(), "wb");
		if (!out) { std::fclose(in); return {}; }
		char buf[4096]; size_t k;
		while ((k = std::fread(buf, 1, sizeof(buf), in)) > 0) std::fwrite(buf, 1, k, out);
		std::fclose(in); std::fclose(out);
		return dst;
	}
};

/* octproz/src/recorder.cpp.  Two ways in:
 *   Recorder(path, bytesPerBuffer, buffersToRecord) + record(buf): N buffers (raw or processed, whatever the caller connects) appended to ONE
 *     headerless little-endian file -- the format the Virtual OCT System replays (docs/docs/faq.md:5); recorder.cpp:99-152 in its plainest form;
 *   Recorder("raw" | "processed") + init(RecordingParams) + record(buf, currentBufferNr) + abort(): the reference's recording session -- file
 *     <savePath>/<timestamp>[_<fileName>]_<name>.raw (:77-82), optional start at the first buffer of a volume (:116-119), capture in memory
 *     and one write when the last buffer has arrived or on abort (:52-62, :124-131). */
class Recorder {
public:
	Recorder(const std::string& path, size_t bytesPerBuffer, unsigned int buffersToRecord)
	    : bytesPerBuffer_(bytesPerBuffer), buffersToRecord_(buffersToRecord), f_(std::fopen(path.c_str(), "wb")), path_(path) {}
	explicit Recorder(const std::string& name) : bytesPerBuffer_(0), buffersToRecord_(0), f_(nullptr), name_(name), session_(true) {}
	Recorder(const Recorder&) = delete;
	Recorder& operator=(const Recorder&) = delete;
	~Recorder() { close(); }
	bool isOpen() const { return f_ != nullptr; }
	unsigned int recordedBuffers() const { return recorded_; }
	bool finished() const { return session_ ? recordingFinished_ : recorded_ >= buffersToRecord_; }
	const std::string& path() const { return path_; }
	/* plain form: returns false once the requested number of buffers has been written (recorder.cpp:124-131: recording finished) */
	bool record(const void* buffer) {
		if (session_) return record(buffer, 0u);
		if (!f_ || finished()) return false;
		if (std::fwrite(buffer, 1, bytesPerBuffer_, f_) != bytesPerBuffer_) { close(); return false; }
		if (++recorded_ == buffersToRecord_) close();
		return true;
	}
	void close() { if (f_) { std::fclose(f_); f_ = nullptr; } }

	/* ---- the reference's session ---- */
	std::function<void(bool)> readyToRecord;           /* signals of recorder.h */
	std::function<void()> recordingDone;
	std::function<void(const std::string&)> error, info;
	bool recordingEnabled() const { return recordingEnabled_; }
	bool init(const RecordingParams& p) {                                                         /* slot_init :64-89 */
		params_ = p;
		std::error_code ec;
		if (p.savePath.empty() || !std::filesystem::is_directory(p.savePath, ec)) {
			say(error, "Recording not initialized: save path is empty or invalid."); uninit(); return false;
		}
		path_ = p.sessionPrefix() + "_" + name_ + ".raw";
		captured_.clear(); captured_.reserve((size_t)p.buffersToRecord * p.bufferSizeInBytes);
		recorded_ = 0; initialized_ = true; recordingFinished_ = false; recordingEnabled_ = true; isRecording_ = false;
		if (readyToRecord) readyToRecord(true);
		say(info, "Recording initialized...");
		return true;
	}
	/* slot_record :99-133; returns true if the buffer was taken */
	bool record(const void* buffer, unsigned int currentBufferNr) {
		if (!recordingEnabled_) return false;
		if (!initialized_) { say(error, "Recording not possible. Record buffer not initialized."); return false; }
		if (params_.startWithFirstBuffer && !isRecording_ && currentBufferNr != 0) return false;
		isRecording_ = true;
		const char* b = static_cast<const char*>(buffer);
		captured_.insert(captured_.end(), b, b + params_.bufferSizeInBytes);
		if (++recorded_ >= params_.buffersToRecord) { recordingEnabled_ = false; isRecording_ = false; saveToDisk(); uninit(); }
		return true;
	}
	void abort() {                                                                                /* slot_abortRecording :52-62 */
		if (recordingEnabled_ && !recordingFinished_) { say(error, "Recording aborted!"); recordingEnabled_ = false; saveToDisk(); uninit(); }
	}

private:
	static void say(const std::function<void(const std::string&)>& f, const std::string& m) { if (f) f(m); }
	void uninit() {
		captured_.clear(); captured_.shrink_to_fit();
		initialized_ = false; recordingFinished_ = true; recorded_ = 0;
		if (readyToRecord) readyToRecord(false);
		if (recordingDone) recordingDone();
	}
	void saveToDisk() {
		if (!initialized_) { say(error, "Save recording to disk not possible. Record buffer not initialized."); return; }
		FILE* out = std::fopen(path_.c_str(), "wb");
		if (!out) { say(error, "Recording failed! Could not write file to disk."); return; }
		say(info, "Captured buffers: " + std::to_string(recorded_) + "/" + std::to_string(params_.buffersToRecord));
		const bool ok = std::fwrite(captured_.data(), 1, captured_.size(), out) == captured_.size();
		std::fclose(out);
		say(ok ? info : error, ok ? "Data written to disk! " + path_ : std::string("Recording failed! Could not write file to disk."));
	}

	size_t bytesPerBuffer_;
	unsigned int buffersToRecord_, recorded_ = 0;
	FILE* f_;
	std::string path_, name_;
	bool session_ = false, recordingEnabled_ = false, recordingFinished_ = false, isRecording_ = false, initialized_ = false;
	RecordingParams params_;
	std::vector<char> captured_;
};

/* processing.cpp:194-207: what the sidebar shows */
struct ProcessingStats {
	double buffersPerSecond = 0, volumesPerSecond = 0, bscansPerSecond = 0, ascansPerSecond = 0, bufferSizeMB = 0, dataThroughputMBs = 0;
	long long processedBuffers = 0;
};

/* Processing::slot_start (processing.cpp:136-229).  Pipeline needs initializeCuda(h1, h2, acq, params*), octCudaPipeline(h), sync(), cleanupCuda(). */
template <class Pipeline>
class Processing {
public:
	Processing(Pipeline* pipeline, OctAlgorithmParameters* octParams) : pipeline_(pipeline), octParams_(octParams) {}

	/* signal rawData(void*, bitDepth, samplesPerLine, ascansPerBscan, bscansPerBuffer, buffersPerVolume, currentBufferNr) (processing.h:110) */
	std::function<void(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned)> rawData;
	ProcessingStats stats;
	Recorder rawRecorder{"raw"}, processedRecorder{"processed"};                              /* processing.cpp:49,60 */
	std::function<void(const std::string&)> error;

	/* Processing::slot_enableRecording (processing.cpp:231-266) with OCTproZApp::slot_prepareGpu2HostForProcessedRecording / slot_resetGpu2HostSettings
	 * (octprozapp.cpp:408-422): raw buffers as this loop sees them and / or the processed buffers the pipeline streams to the host (converted
	 * containers, or float32 with saveAs32bitFloat); the meta file is a copy of `settingsFile`.  Call before slot_start or from the rawData slot.
	 * The pipeline type needs cuda_register[Float]StreamingBuffers / cuda_unregister... and currentBufferNr() (OctPipeline has them). */
	void slot_enableRecording(const RecordingParams& rp, const AcquisitionParams& a, const std::string& settingsFile = std::string()) {
		if (rp.recordRaw) {
			if (rawRecorder.recordingEnabled()) say("Recording of raw data is already running.");
			else rawRecorder.init(rp);
		}
		if (rp.recordProcessed) {
			if (processedRecorder.recordingEnabled()) say("Recording of processed data is already running.");
			else {
				octb200_params& p = octParams_->p;
				memorized_[0] = p.streamToHost; memorized_[1] = (int)p.streamingBuffersToSkip; memorized_[2] = p.streamFloatToHost; haveMemorized_ = true;
				p.streamToHost = 1; p.streamingBuffersToSkip = 0; p.streamFloatToHost = rp.saveAs32bitFloat ? 1 : 0;     /* every buffer, while the recording runs */
				processedRecorder.recordingDone = [this] {
					if (!haveMemorized_) return;
					octb200_params& q = octParams_->p;
					q.streamToHost = memorized_[0]; q.streamingBuffersToSkip = (uint32_t)memorized_[1]; q.streamFloatToHost = memorized_[2]; haveMemorized_ = false;
				};
				if (processedRecorder.init(rp.forProcessedData(a))) {
					streamingWanted_ = true; recordFloat_ = rp.saveAs32bitFloat; streamBytes_ = rp.forProcessedData(a).bufferSizeInBytes;
					/* bound here, not in slot_start: only pipelines that record processed data need the streaming half of kernels.h */
					enableStreaming_ = [this] { enableStreaming(); };
					disableStreaming_ = [this] { disableStreaming(); };
				}
			}
		}
		if (!settingsFile.empty()) rp.saveMeta(settingsFile);
	}

	/* maxBuffers > 0: headless runs stop the acquisition after that many processed buffers (the GUI's Stop button) */
	bool slot_start(AcquisitionSystem* system, long long maxBuffers = 0) {
		AcquisitionBuffer* buffer = system->buffer;
		for (int i = 0; i < buffer->bufferCnt; ++i) buffer->setReady(i, true);                  /* blockBuffersForAcquisitionSystem :124-128 */
		const AcquisitionParams& a = system->params;
		if (buffer->bufferCnt < 2 || !pipeline_->initializeCuda(buffer->bufferArray[0], buffer->bufferArray[1], a, octParams_)) {      /* :151 */
			for (int i = 0; i < buffer->bufferCnt; ++i) buffer->setReady(i, false);
			system->stopAcquisition();                                                            /* initializationFailed -> slot_stop (octprozapp.cpp:54) */
			return false;
		}
		const unsigned perVolume = a.buffersPerVolume ? a.buffersPerVolume : 1;
		unsigned currentBufferNr = perVolume - 1;
		for (int i = 0; i < buffer->bufferCnt; ++i) buffer->setReady(i, false);                 /* unblock :130-134 */
		const auto t0 = std::chrono::steady_clock::now();
		long long n = 0;
		while (system->acqusitionRunning.load()) {                                               /* :176-218 */
			const int pos = buffer->currIndex.load();
			if (pos >= 0 && buffer->ready(pos)) {
				currentBufferNr = (currentBufferNr + 1) % perVolume;
				if (streamingWanted_ && enableStreaming_) enableStreaming_();
				if (rawData) rawData(buffer->bufferArray[pos], a.bitDepth, a.samplesPerLine, a.ascansPerBscan, a.bscansPerBuffer, perVolume, currentBufferNr);
				rawRecorder.record(buffer->bufferArray[pos], currentBufferNr);                      /* connect(rawData, rawRecorder) :52; a no-op unless enabled */
				pipeline_->octCudaPipeline(buffer->bufferArray[pos]);                           /* :187 */
				buffer->setReady(pos, false);                                                    /* :191 */
				++n;
				if (maxBuffers > 0 && n >= maxBuffers) system->stopAcquisition();
			} else {
				std::this_thread::yield();
			}
		}
		pipeline_->sync();
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (disableStreaming_) disableStreaming_();
		const double bps = dt > 0 ? (double)n / dt : 0.0;
		stats.processedBuffers = n;
		stats.buffersPerSecond = bps; stats.volumesPerSecond = bps / perVolume;                 /* :198-201 */
		stats.bscansPerSecond = bps * a.bscansPerBuffer; stats.ascansPerSecond = stats.bscansPerSecond * a.ascansPerBscan;
		stats.bufferSizeMB = (double)buffer->bytesPerBuffer / 1048576.0; stats.dataThroughputMBs = bps * stats.bufferSizeMB;
		return true;
	}

private:
	void say(const std::string& m) { if (error) error(m); }
	/* enableGpu2HostStreaming / enableFloatGpu2HostStreaming (processing.cpp:316-362): two host buffers registered with the pipeline, every
	 * delivered buffer handed to the processed recorder through the notifier */
	void enableStreaming() {
		streamingWanted_ = false;
		if (!streamBuffer_.allocateMemory(2, streamBytes_)) { say("could not allocate the streaming buffers"); return; }
		auto deliver = [this](void* b) { processedRecorder.record(b, pipeline_->currentBufferNr()); };
		if (recordFloat_) {
			Gpu2HostNotifier::getInstance().connectFloat(deliver);
			pipeline_->cuda_registerFloatStreamingBuffers(streamBuffer_.bufferArray[0], streamBuffer_.bufferArray[1], streamBytes_);
		} else {
			Gpu2HostNotifier::getInstance().connectProcessed(deliver);
			pipeline_->cuda_registerStreamingBuffers(streamBuffer_.bufferArray[0], streamBuffer_.bufferArray[1], streamBytes_);
		}
		streaming_ = true;
	}
	void disableStreaming() {
		if (!streaming_) return;
		if (recordFloat_) { pipeline_->cuda_unregisterFloatStreamingBuffers(); Gpu2HostNotifier::getInstance().connectFloat(nullptr); }
		else { pipeline_->cuda_unregisterStreamingBuffers(); Gpu2HostNotifier::getInstance().connectProcessed(nullptr); }
		streamBuffer_.releaseMemory();
		streaming_ = false;
	}

	Pipeline* pipeline_;
	OctAlgorithmParameters* octParams_;
	AcquisitionBuffer streamBuffer_;
	std::function<void()> enableStreaming_, disableStreaming_;
	size_t streamBytes_ = 0;
	bool streamingWanted_ = false, streaming_ = false, recordFloat_ = false, haveMemorized_ = false;
	int memorized_[3] = {0, 0, 0};
};

/* run `buffers` buffers of a raw file through `pipeline` with the reference's thread structure (acquisition thread + processing loop) */
template <class Pipeline>
ProcessingStats replay(VirtualOCTSystem& vos, Pipeline& pipeline, OctAlgorithmParameters& params, long long buffers, bool* ok = nullptr) {
	Processing<Pipeline> proc(&pipeline, &params);
	std::atomic<int> state{0};          /* 1 = started, 2 = stopped before starting (file missing) */
	vos.acquisitionStarted = [&](AcquisitionSystem*) { state.store(1); };
	vos.acquisitionStopped = [&]() { int expected = 0; state.compare_exchange_strong(expected, 2); };
	std::thread producer([&] { vos.startAcquisition(); });
	while (state.load() == 0) std::this_thread::yield();
	bool good = false;
	if (state.load() == 1) good = proc.slot_start(&vos, buffers);
	vos.stopAcquisition();
	producer.join();
	if (ok) *ok = good;
	return proc.stats;
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * Dispersion estimator (octproz-dispersion-estimator-extension): DispersionEstimationEngine::startDispersionEstimation
 * (src/dispersionestimationengine.cpp:21-116) with the reference's search semantics -- d2 sweep with d3 = 0, then d3 at the best d2,
 * strict '<' from a best metric of 0, step |end - start| / n, the two plotted A-scans, d1 = -(d2 + d3).  The reference re-runs its
 * CPU path once per trial value (processDispersionMetric, :118-158); here a sweep is ONE call of `Sweep` -- for the product
 * PipelineSweep, i.e. octb200_dispersion_sweep: all trials in one launch of the fused kernel plus one metric kernel.
 * Sweep: std::vector<float> operator()(const void* raw, unsigned lines, const std::vector<float>& coeffs, std::vector<float>* ascans)
 *        coeffs = four floats per trial (d0 d1 d2 d3), ascans = NULL or [trials][lines][N/2], returns one metric per trial
 * ------------------------------------------------------------------------------------------------------------------------------- */
struct DispersionEstimatorParameters {                                  /* src/dispersionestimatorparameters.h:58-76 (GUI-only fields omitted) */
	int numberOfCenterAscans = 10;
	bool useLinearAscans = true;
	int numberOfAscanSamplesToIgnore = 0;
	bool autoCalcD1 = false;
	int sharpnessMetric = OCTB200_METRIC_SUM_ABOVE_THRESHOLD;
	float metricThreshold = 0.0f;
	double d2start = -100.0, d2end = 100.0, d3start = -100.0, d3end = 100.0;
	int numberOfDispersionSamples = 100;
};

/* the window the estimator's processing path applies whatever the main window setting is (octprocessor/processor.tpp:126-133, T = float) */
inline std::vector<float> cpuPathWindow(unsigned int samplesPerSpectrum) {
	std::vector<float> w(samplesPerSpectrum);
	const float factor = static_cast<float>(2.0 * 3.14159265358979323846 / (samplesPerSpectrum - 1));
	for (size_t i = 0; i < samplesPerSpectrum; ++i) w[i] = static_cast<float>(0.5) * (1 - std::cos(factor * i));
	return w;
}

template <class Sweep>
class DispersionEstimationEngine {
public:
	DispersionEstimatorParameters params;
	double bestD2 = 0, bestD3 = 0, bestMetricValueD2 = 0, bestMetricValueD3 = 0, calculatedD1 = 0;
	std::vector<std::pair<double, float>> metricsD2, metricsD3;         /* (trial value, metric): what the extension plots */
	std::vector<float> ascanWithoutDispersionCompensation, ascanWithBestDispersion;

	/* d0, d1: the main settings' dispersion coefficients, kept while d2 / d3 are searched (processorcontroller.cpp:59-66) */
	DispersionEstimationEngine(Sweep sweep, float d0, float d1) : sweep_(sweep), d0_(d0), d1_(d1) {}
	void setParams(const DispersionEstimatorParameters& p) { params = p; }

	void startDispersionEstimation(const void* frameBuffer, unsigned int bitDepth, unsigned int samplesPerLine, unsigned int linesPerFrame) {
		const unsigned int centerAscans = std::min(static_cast<unsigned int>(params.numberOfCenterAscans), linesPerFrame);      /* :37 */
		unsigned int offsetAscans = 0;
		if (centerAscans < linesPerFrame) offsetAscans = (linesPerFrame - centerAscans) / 2;                                   /* :42-47 */
		const size_t bytesPerSample = static_cast<size_t>(std::ceil(static_cast<double>(bitDepth) / 8.0));
		const size_t lineSizeBytes = samplesPerLine * bytesPerSample;
		const char* rawData = static_cast<const char*>(frameBuffer) + offsetAscans * lineSizeBytes;                            /* :50-57 */

		const int n = params.numberOfDispersionSamples;
		const double stepSizeD2 = std::fabs(params.d2end - params.d2start) / static_cast<double>(n);                           /* :69 */
		const double stepSizeD3 = std::fabs(params.d3end - params.d3start) / static_cast<double>(n);                           /* :70 */

		/* d2, with d3 = 0 (:78-82) */
		std::vector<double> trial(n);
		{ double d = params.d2start; for (int i = 0; i < n; ++i) { trial[i] = d; d += stepSizeD2; } }
		std::vector<float> m = sweep_(rawData, centerAscans, coeffs(trial, true, 0.0), nullptr);
		bestD2 = 0; bestMetricValueD2 = 0; metricsD2.clear();
		for (int i = 0; i < n; ++i) {
			if (bestMetricValueD2 < m[i]) { bestMetricValueD2 = m[i]; bestD2 = trial[i]; }                                     /* :140-143, strict */
			metricsD2.emplace_back(trial[i], m[i]);
		}
		/* d3, at the best d2 (:85-90) */
		{ double d = params.d3start; for (int i = 0; i < n; ++i) { trial[i] = d; d += stepSizeD3; } }
		m = sweep_(rawData, centerAscans, coeffs(trial, false, bestD2), nullptr);
		bestD3 = 0; bestMetricValueD3 = 0; metricsD3.clear();
		for (int i = 0; i < n; ++i) {
			if (bestMetricValueD3 < m[i]) { bestMetricValueD3 = m[i]; bestD3 = trial[i]; }                                     /* :147-150 */
			metricsD3.emplace_back(trial[i], m[i]);
		}
		/* the two A-scans the GUI plots (:93-96; processFirstLineOnly :160-192 applies the center offset a second time inside the
		   already extracted block -- with centerAscans lines in the block that offset is 0) */
		std::vector<float> ascans;
		const std::vector<float> two = { d0_, d1_, 0.0f, 0.0f, d0_, d1_, static_cast<float>(bestD2), static_cast<float>(bestD3) };
		sweep_(rawData, 1u, two, &ascans);
		const size_t half = samplesPerLine / 2;
		ascanWithoutDispersionCompensation.assign(ascans.begin(), ascans.begin() + half);
		ascanWithBestDispersion.assign(ascans.begin() + half, ascans.begin() + 2 * half);
		if (params.autoCalcD1) calculatedD1 = -(bestD2 + bestD3);                                                              /* :101 */
	}

private:
	std::vector<float> coeffs(const std::vector<double>& trial, bool isD2, double other) const {
		std::vector<float> c;
		c.reserve(trial.size() * 4);
		for (double t : trial) {
			c.push_back(d0_); c.push_back(d1_);
			c.push_back(static_cast<float>(isD2 ? t : other)); c.push_back(static_cast<float>(isD2 ? other : t));
		}
		return c;
	}
	Sweep sweep_;
	float d0_, d1_;
};

/* The product sweep: octb200_dispersion_sweep on an initialised OctPipeline.  The estimator's processing path differs from the main one
 * in three settings, applied for the lifetime of this object and restored afterwards: its own Hanning window (cpuPathWindow), no
 * bitshift, and -- a quirk of the reference reproduced on purpose -- a rolling-average window of 10 whatever the settings say
 * (processorcontroller.cpp:116 passes the setting into the constructor's `windowSize` slot, processor.h:26). */
class PipelineSweep {
public:
	PipelineSweep(OctPipeline& pipe, OctAlgorithmParameters& q, unsigned int samplesPerLine, const DispersionEstimatorParameters& prm)
	    : pipe_(pipe), q_(q), n_(samplesPerLine), prm_(prm), savedParams_(q.p), savedWindow_(q.windowCurve) {
		q_.p.bitshift = 0;
		q_.p.rollingAverageWindowSize = 10;
		q_.windowCurve = cpuPathWindow(samplesPerLine);
		apply();
	}
	~PipelineSweep() {
		q_.p = savedParams_;
		q_.windowCurve = savedWindow_;
		try { apply(); } catch (...) {}
	}
	std::vector<float> operator()(const void* raw, unsigned int lines, const std::vector<float>& coeffs, std::vector<float>* ascans) {
		octb200_sweep_config c;
		std::memset(&c, 0, sizeof(c));
		c.lines = lines; c.trials = static_cast<uint32_t>(coeffs.size() / 4); c.metric = prm_.sharpnessMetric; c.metricThreshold = prm_.metricThreshold;
		c.samplesToIgnore = prm_.numberOfAscanSamplesToIgnore; c.logScale = prm_.useLinearAscans ? 0 : 1;
		c.logMin = q_.p.signalGrayscaleMin; c.logMax = q_.p.signalGrayscaleMax; c.logCoeff = q_.p.signalMultiplicator; c.logAddend = q_.p.signalAddend;
		std::vector<float> metrics(c.trials);
		if (ascans) ascans->assign(static_cast<size_t>(c.trials) * lines * (n_ / 2), 0.0f);
		const int rc = octb200_dispersion_sweep(pipe_.handle(), raw, &c, coeffs.data(), metrics.data(), ascans ? ascans->data() : nullptr);
		if (rc != OCTB200_OK) throw std::runtime_error(std::string("octb200_dispersion_sweep failed: ") + octb200_last_error(pipe_.handle()));
		return metrics;
	}

private:
	void apply() {
		if (octb200_set_params(pipe_.handle(), &q_.p) != OCTB200_OK) throw std::runtime_error("set_params failed");
		if (q_.p.windowing && octb200_set_window_curve(pipe_.handle(), q_.windowCurve.data(), static_cast<int>(n_)) != OCTB200_OK)
			throw std::runtime_error("set_window_curve failed");
	}
	OctPipeline& pipe_;
	OctAlgorithmParameters& q_;
	unsigned int n_;
	DispersionEstimatorParameters prm_;
	octb200_params savedParams_;
	std::vector<float> savedWindow_;
};

}  // namespace host
}  // namespace octb200

#endif /* OCTB200_HOST_HPP */
