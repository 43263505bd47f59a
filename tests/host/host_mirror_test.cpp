/* CPU test of include/octb200_host.hpp: the acquisition / processing handshake of the reference (processing.cpp:124-229,
 * virtualoctsystem.cpp:143-224) against a stand-in pipeline -- no GPU, no CUDA call.  TEST ONLY: the stand-in computes nothing.
 * Built and run by tests/test_host_mirror.py; exit code 0 = all checks passed. */
#include <cstdio>
#include <string>
#include <vector>

#include "octb200_host.hpp"

using namespace octb200::host;

#define CHECK(cond)                                                                      \
	do {                                                                                 \
		if (!(cond)) { std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); return 1; } \
	} while (0)

struct FakePipeline {
	bool failInit = false;
	void* h1 = nullptr; void* h2 = nullptr;
	AcquisitionParams acq;
	std::vector<void*> calls;
	std::vector<unsigned short> firstSample;
	int syncs = 0, cleanups = 0;
	bool initializeCuda(void* a, void* b, const AcquisitionParams& p, OctAlgorithmParameters*) { h1 = a; h2 = b; acq = p; return !failInit; }
	void octCudaPipeline(void* h) { calls.push_back(h); firstSample.push_back(*static_cast<unsigned short*>(h)); }
	void sync() { ++syncs; }
	void cleanupCuda() { ++cleanups; }
};

/* stand-in with the streaming half of kernels.h: the "processed" buffer is the first half of the raw buffer + 1 (u16) or the same as float,
 * written into the registered host buffers in turn and announced through the notifier's C callbacks, as the library does */
struct FakeStreamingPipeline : FakePipeline {
	OctAlgorithmParameters* q = nullptr;
	void* stream[2] = {nullptr, nullptr}; void* fstream[2] = {nullptr, nullptr};
	size_t bytes = 0, fbytes = 0;
	unsigned nr = 0, perVolume = 1, count = 0;
	int registered = 0, unregistered = 0;
	bool initializeCuda(void* a, void* b, const AcquisitionParams& p, OctAlgorithmParameters* params) {
		q = params; perVolume = p.buffersPerVolume ? p.buffersPerVolume : 1; nr = perVolume - 1;
		return FakePipeline::initializeCuda(a, b, p, params);
	}
	void cuda_registerStreamingBuffers(void* h1, void* h2, size_t n) { stream[0] = h1; stream[1] = h2; bytes = n; ++registered; }
	void cuda_unregisterStreamingBuffers() { stream[0] = stream[1] = nullptr; ++unregistered; }
	void cuda_registerFloatStreamingBuffers(void* h1, void* h2, size_t n) { fstream[0] = h1; fstream[1] = h2; fbytes = n; ++registered; }
	void cuda_unregisterFloatStreamingBuffers() { fstream[0] = fstream[1] = nullptr; ++unregistered; }
	unsigned currentBufferNr() const { return nr; }
	void octCudaPipeline(void* h) {
		FakePipeline::octCudaPipeline(h);
		nr = (nr + 1) % perVolume;
		const unsigned short* raw = static_cast<const unsigned short*>(h);
		if (q->p.streamToHost && stream[0] && !q->p.streamFloatToHost) {
			unsigned short* dst = static_cast<unsigned short*>(stream[count & 1]);
			for (size_t i = 0; i < bytes / 2; ++i) dst[i] = (unsigned short)(raw[i] + 1);
			Gpu2HostNotifier::dh2StreamingCallback(dst);
		}
		if (q->p.streamFloatToHost && fstream[0]) {
			float* dst = static_cast<float*>(fstream[count & 1]);
			for (size_t i = 0; i < fbytes / 4; ++i) dst[i] = (float)(raw[i] + 1);
			Gpu2HostNotifier::dh2FloatStreamingCallback(dst);
		}
		++count;
	}
};

static bool writeFile(const std::string& path, unsigned n, unsigned a, unsigned b, unsigned buffers) {
	FILE* f = std::fopen(path.c_str(), "wb");
	if (!f) return false;
	std::vector<unsigned short> buf((size_t)n * a * b);
	for (unsigned k = 0; k < buffers; ++k) {
		for (size_t i = 0; i < buf.size(); ++i) buf[i] = (unsigned short)((1000u * (k + 1) + i) & 0xFFFu);
		std::fwrite(buf.data(), 2, buf.size(), f);
	}
	std::fclose(f);
	return true;
}

int main(int argc, char** argv) {
	const std::string path = argc > 1 ? argv[1] : "/tmp/octb200_host_mirror.raw";
	const unsigned n = 64, a = 8, b = 2;
	CHECK(writeFile(path, n, a, b, 3));

	/* AcquisitionBuffer: 128-byte alignment, zeroed, flags down (acquisitionbuffer.cpp:43-63) */
	{
		AcquisitionBuffer ab;
		CHECK(ab.allocateMemory(2, 1000));
		CHECK(ab.bufferCnt == 2 && ab.bytesPerBuffer == 1000 && ab.currIndex.load() == -1);
		for (int i = 0; i < 2; ++i) { CHECK(((uintptr_t)ab.bufferArray[i] & 127u) == 0); CHECK(!ab.ready(i)); CHECK(static_cast<unsigned char*>(ab.bufferArray[i])[999] == 0); }
		ab.releaseMemory();
		CHECK(ab.bufferCnt == 0 && ab.bufferArray.empty());
	}

	/* replay: two buffers from the file, alternating, every delivered buffer processed exactly once while synchronised */
	{
		VirtualOCTSystem vos(path, 12, n, a, b, 1);
		FakePipeline fp;
		OctAlgorithmParameters q = OctAlgorithmParameters();
		bool ok = false;
		ProcessingStats st = replay(vos, fp, q, 7, &ok);
		CHECK(ok && st.processedBuffers == 7 && fp.calls.size() == 7 && fp.syncs == 1);
		CHECK(fp.h1 == vos.buffer->bufferArray[0] && fp.h2 == vos.buffer->bufferArray[1]);
		CHECK(fp.acq.samplesPerLine == n && fp.acq.ascansPerBscan == a && fp.acq.bscansPerBuffer == b && fp.acq.bitDepth == 12);
		for (size_t i = 1; i < fp.calls.size(); ++i) CHECK(fp.calls[i] != fp.calls[i - 1]);               /* strict alternation of the two slots */
		for (size_t i = 0; i < fp.calls.size(); ++i) {
			const unsigned short want = (unsigned short)((fp.calls[i] == fp.h1 ? 1000u : 2000u) & 0xFFFu);   /* buffer 0 / 1 of the file */
			CHECK(fp.firstSample[i] == want);
		}
		CHECK(st.ascansPerSecond > 0 && st.bufferSizeMB > 0);
		CHECK(vos.buffer->bytesPerBuffer == (size_t)n * a * b * 2);
	}

	/* buffersFromFile = 1 replays the same buffer into both slots; bscanOffset skips B-scans (virtualoctsystem.cpp:167-179) */
	{
		VirtualOCTSystem vos(path, 12, n, a, b, 1);
		vos.buffersFromFile = 1; vos.bscanOffset = b;          /* start at the second buffer of the file */
		FakePipeline fp;
		OctAlgorithmParameters q;
		bool ok = false;
		replay(vos, fp, q, 4, &ok);
		CHECK(ok && fp.calls.size() == 4);
		for (size_t i = 0; i < fp.calls.size(); ++i) CHECK(fp.firstSample[i] == (unsigned short)(2000u & 0xFFFu));
	}

	/* buffersFromFile > 2 streams the file buffer by buffer and rewinds after buffersFromFile buffers (virtualoctsystem.cpp:226-290) */
	{
		VirtualOCTSystem vos(path, 12, n, a, b, 1);
		vos.buffersFromFile = 3;
		FakePipeline fp;
		OctAlgorithmParameters q;
		bool ok = false;
		replay(vos, fp, q, 8, &ok);
		CHECK(ok && fp.calls.size() == 8);
		/* consecutive buffers of the file, cyclic.  Where the cycle starts depends on the start-up handshake: Processing raises and then
		   clears ALL ready flags around initializeCuda (processing.cpp:124-134), which can discard a buffer the acquisition thread had
		   already delivered -- in the reference as well */
		unsigned off = 0;
		for (unsigned k = 0; k < 3; ++k) if (fp.firstSample[0] == (unsigned short)((1000u * (k + 1)) & 0xFFFu)) off = k;
		for (size_t i = 0; i < fp.calls.size(); ++i) CHECK(fp.firstSample[i] == (unsigned short)((1000u * (unsigned)((i + off) % 3 + 1)) & 0xFFFu));
		for (size_t i = 1; i < fp.calls.size(); ++i) CHECK(fp.calls[i] != fp.calls[i - 1]);
	}

	/* failed initialisation stops the acquisition and releases the flags (processing.cpp:151-160) */
	{
		VirtualOCTSystem vos(path, 12, n, a, b, 1);
		FakePipeline fp; fp.failInit = true;
		OctAlgorithmParameters q;
		bool ok = true;
		ProcessingStats st = replay(vos, fp, q, 3, &ok);
		CHECK(!ok && st.processedBuffers == 0 && fp.calls.empty());
		CHECK(!vos.acqusitionRunning.load());
	}

	/* a missing file never starts (virtualoctsystem.cpp:112-118) */
	{
		VirtualOCTSystem vos(path + ".missing", 12, n, a, b, 1);
		FakePipeline fp;
		OctAlgorithmParameters q;
		bool ok = true;
		replay(vos, fp, q, 3, &ok);
		CHECK(!ok && fp.calls.empty());
	}

	/* the benchmark parameter block and its curves (host generators of liboctb200, bit-exact against the reference's: tests/test_curves.py) */
	{
		OctAlgorithmParameters q = OctAlgorithmParameters::benchmark(1024);
		CHECK(q.resampleCurve.size() == 1024 && q.windowCurve.size() == 1024 && q.dispersionCurve.size() == 1024);
		CHECK(q.p.resampling && q.p.windowing && q.p.dispersionCompensation && q.p.fixedPatternNoiseRemoval && q.p.signalLogScaling);
		CHECK(q.windowCurve[26] == 0.0f && q.windowCurve[512] > 0.9999f);          /* SURVEY 8c known answers: Hann 0.95 / 0.5 */
		CHECK(q.resampleCurve[0] >= 0.0f && q.resampleCurve[1023] <= 1021.0f);     /* Polynomial::clamp to [0, N - 3] */
	}

	/* the product pipeline type has no CPU fallback: without a GPU initializeCuda fails loudly, with one it works */
	{
		OctPipeline p;
		AcquisitionParams acq; acq.samplesPerLine = 7; acq.ascansPerBscan = 4; acq.bscansPerBuffer = 1; acq.buffersPerVolume = 1; acq.bitDepth = 12;
		OctAlgorithmParameters q;
		CHECK(!p.initializeCuda(nullptr, nullptr, acq, &q));
		CHECK(p.lastError().find("geometry") != std::string::npos);
	}
	/* Recorder: the raw buffers seen by Processing (signal rawData, processing.h:110) recorded into one headerless file that replays identically */
	{
		const std::string rec = path + ".rec";
		VirtualOCTSystem vos(path, 12, n, a, b, 1);
		FakePipeline fp;
		OctAlgorithmParameters q;
		Recorder recorder(rec, (size_t)n * a * b * 2, 3);
		CHECK(recorder.isOpen());
		Processing<FakePipeline> proc(&fp, &q);
		proc.rawData = [&](void* buf, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned) { recorder.record(buf); };
		std::atomic<bool> started{false};
		vos.acquisitionStarted = [&](AcquisitionSystem*) { started.store(true); };
		std::thread producer([&] { vos.startAcquisition(); });
		while (!started.load()) std::this_thread::yield();
		CHECK(proc.slot_start(&vos, 5));
		vos.stopAcquisition();
		producer.join();
		CHECK(recorder.finished() && recorder.recordedBuffers() == 3 && !recorder.record(vos.buffer->bufferArray[0]));
		FILE* f = std::fopen(rec.c_str(), "rb");
		CHECK(f);
		std::vector<unsigned short> back((size_t)n * a * b * 3);
		CHECK(std::fread(back.data(), 2, back.size(), f) == back.size() && std::fgetc(f) == EOF);
		std::fclose(f);
		for (int k = 0; k < 3; ++k) CHECK(back[(size_t)k * n * a * b] == fp.firstSample[k]);                     /* the buffers that were processed, in order */
	}

	/* Recorder, session form (recorder.cpp:64-152): file name rule, start with the first buffer of a volume, completion, abort, bad save path, meta file */
	{
		const std::string dir = path.substr(0, path.find_last_of('/'));
		RecordingParams rp; rp.timestamp = "20261017_101500"; rp.fileName = "phantom"; rp.savePath = dir; rp.bufferSizeInBytes = 16; rp.buffersToRecord = 3;
		rp.startWithFirstBuffer = true; rp.saveMetaData = true;
		std::vector<std::string> infos, errors; int done = 0;
		Recorder r("raw");
		r.info = [&](const std::string& m) { infos.push_back(m); }; r.error = [&](const std::string& m) { errors.push_back(m); }; r.recordingDone = [&] { ++done; };
		unsigned short x[4][8];
		for (int k = 0; k < 4; ++k) for (int i = 0; i < 8; ++i) x[k][i] = (unsigned short)(100 * k + i);
		CHECK(!r.record(x[0], 0u));                                                  /* not enabled yet */
		CHECK(r.init(rp) && r.path() == dir + "/20261017_101500_phantom_raw.raw" && r.recordingEnabled());
		CHECK(!r.record(x[3], 1u) && r.recordedBuffers() == 0);                      /* waits for the first buffer of a volume */
		CHECK(r.record(x[0], 0u) && r.record(x[1], 1u) && r.record(x[2], 0u) && !r.record(x[3], 1u));
		CHECK(done == 1 && r.finished() && !r.recordingEnabled());
		FILE* f = std::fopen(r.path().c_str(), "rb");
		CHECK(f);
		unsigned short back[24];
		CHECK(std::fread(back, 2, 24, f) == 24 && std::fgetc(f) == EOF);
		std::fclose(f);
		for (int k = 0; k < 3; ++k) for (int i = 0; i < 8; ++i) CHECK(back[8 * k + i] == x[k][i]);
		CHECK(!infos.empty() && infos[1] == "Captured buffers: 3/3");
		/* processed-data sizes (processing.cpp:243-249) */
		AcquisitionParams acq; acq.samplesPerLine = 64; acq.ascansPerBscan = 4; acq.bscansPerBuffer = 3; acq.buffersPerVolume = 1; acq.bitDepth = 12;
		RecordingParams raw = rp; raw.bufferSizeInBytes = 64 * 4 * 3 * 2;
		CHECK(raw.forProcessedData(acq).bufferSizeInBytes == 64 * 4 * 3);
		raw.saveAs32bitFloat = true;
		CHECK(raw.forProcessedData(acq).bufferSizeInBytes == 32 * 4 * 3 * 4);
		/* abort keeps what was captured; no user file name -> no extra underscore */
		RecordingParams rp2 = rp; rp2.fileName.clear(); rp2.timestamp = "t"; rp2.buffersToRecord = 5; rp2.startWithFirstBuffer = false;
		Recorder r2("processed");
		r2.error = [&](const std::string& m) { errors.push_back(m); };
		CHECK(r2.init(rp2) && r2.path() == dir + "/t_processed.raw");
		CHECK(r2.record(x[0], 1u) && r2.record(x[3], 0u));
		r2.abort(); r2.abort();
		CHECK(errors.size() == 1 && errors[0] == "Recording aborted!" && r2.finished());
		f = std::fopen(r2.path().c_str(), "rb");
		CHECK(f && std::fread(back, 2, 24, f) == 16 && back[8] == x[3][0]);
		std::fclose(f);
		/* meta file = copy of the settings file */
		const std::string ini = dir + "/settings_for_meta.ini";
		f = std::fopen(ini.c_str(), "wb"); CHECK(f); std::fputs("[processing]\nbitshift=false\n", f); std::fclose(f);
		const std::string meta = rp.saveMeta(ini);
		CHECK(meta == dir + "/20261017_101500_phantom_meta.txt");
		f = std::fopen(meta.c_str(), "rb"); CHECK(f);
		char txt[64] = {}; CHECK(std::fread(txt, 1, 63, f) > 0 && std::string(txt) == "[processing]\nbitshift=false\n"); std::fclose(f);
		/* invalid save path */
		RecordingParams bad = rp; bad.savePath = dir + "/does_not_exist";
		Recorder r3("raw");
		r3.error = [&](const std::string& m) { errors.push_back(m); };
		CHECK(!r3.init(bad) && !r3.recordingEnabled() && errors.back().find("save path") != std::string::npos);
	}

	/* Processing::slot_enableRecording (processing.cpp:231-266): raw + processed session over the replay loop, containers and float32 */
	for (int asFloat = 0; asFloat < 2; ++asFloat) {
		const std::string dir = path.substr(0, path.find_last_of('/'));
		VirtualOCTSystem vos(path, 12, n, a, b, 2);                                /* two buffers per volume */
		FakeStreamingPipeline fp;
		OctAlgorithmParameters q;
		q.p.streamToHost = 0; q.p.streamingBuffersToSkip = 3;
		Processing<FakeStreamingPipeline> proc(&fp, &q);
		std::vector<std::string> errors;
		proc.error = [&](const std::string& m) { errors.push_back(m); };
		RecordingParams rp; rp.timestamp = asFloat ? "tsf" : "ts"; rp.savePath = dir; rp.bufferSizeInBytes = (size_t)n * a * b * 2; rp.buffersToRecord = 4;
		rp.startWithFirstBuffer = true; rp.recordRaw = true; rp.recordProcessed = true; rp.saveAs32bitFloat = asFloat != 0;
		proc.slot_enableRecording(rp, vos.params);
		CHECK(q.p.streamToHost == 1 && q.p.streamingBuffersToSkip == 0 && q.p.streamFloatToHost == asFloat);
		proc.slot_enableRecording(rp, vos.params);                                 /* a second request while one is running */
		CHECK(errors.size() == 2 && errors[0] == "Recording of raw data is already running." && errors[1] == "Recording of processed data is already running.");
		std::atomic<bool> started{false};
		vos.acquisitionStarted = [&](AcquisitionSystem*) { started.store(true); };
		std::thread producer([&] { vos.startAcquisition(); });
		while (!started.load()) std::this_thread::yield();
		CHECK(proc.slot_start(&vos, 9));
		vos.stopAcquisition();
		producer.join();
		CHECK(proc.rawRecorder.finished() && proc.processedRecorder.finished());
		CHECK(q.p.streamToHost == 0 && q.p.streamingBuffersToSkip == 3 && q.p.streamFloatToHost == 0);        /* settings restored (octprozapp.cpp:418-422) */
		CHECK(fp.registered == 1 && fp.unregistered == 1 && !fp.stream[0] && !fp.fstream[0]);
		const size_t perBuf = (size_t)n * a * b;
		std::vector<unsigned short> raw(perBuf * 4);
		FILE* f = std::fopen((dir + "/" + rp.timestamp + "_raw.raw").c_str(), "rb");
		CHECK(f && std::fread(raw.data(), 2, raw.size(), f) == raw.size() && std::fgetc(f) == EOF);
		std::fclose(f);
		/* four consecutive buffers of the file's two (first samples 1000 / 2000 & 0xFFF), alternating */
		const unsigned short s0 = (unsigned short)(1000u & 0xFFFu), s1 = (unsigned short)(2000u & 0xFFFu);
		CHECK(raw[0] == s0 || raw[0] == s1);
		for (int k = 1; k < 4; ++k) CHECK(raw[k * perBuf] == (raw[(k - 1) * perBuf] == s0 ? s1 : s0));
		f = std::fopen((dir + "/" + rp.timestamp + "_processed.raw").c_str(), "rb");
		CHECK(f);
		if (asFloat) {
			std::vector<float> pr(perBuf / 2 * 4);
			CHECK(std::fread(pr.data(), 4, pr.size(), f) == pr.size() && std::fgetc(f) == EOF);
			for (int k = 0; k < 4; ++k) CHECK(pr[k * (perBuf / 2)] == (float)(s0 + 1) || pr[k * (perBuf / 2)] == (float)(s1 + 1));
			for (int k = 1; k < 4; ++k) CHECK(pr[k * (perBuf / 2)] != pr[(k - 1) * (perBuf / 2)]);
			CHECK(pr[1] == pr[0] + 1.0f);                                          /* the file's ramp, + 1 */
		} else {
			std::vector<unsigned short> pr(perBuf / 2 * 4);
			CHECK(std::fread(pr.data(), 2, pr.size(), f) == pr.size() && std::fgetc(f) == EOF);              /* half the raw buffer's bytes (processing.cpp:248) */
			for (int k = 0; k < 4; ++k) CHECK(pr[k * (perBuf / 2)] == s0 + 1 || pr[k * (perBuf / 2)] == s1 + 1);
			for (int k = 1; k < 4; ++k) CHECK(pr[k * (perBuf / 2)] != pr[(k - 1) * (perBuf / 2)]);
		}
		std::fclose(f);
	}

	/* dispersion estimator search (dispersionestimationengine.cpp:21-116) against a synthetic metric with a known optimum */
	{
		struct FakeSweep {
			int* calls; std::vector<unsigned>* linesSeen; const void** rawSeen;
			std::vector<float> operator()(const void* raw, unsigned lines, const std::vector<float>& c, std::vector<float>* ascans) {
				++*calls; linesSeen->push_back(lines); *rawSeen = raw;
				std::vector<float> m(c.size() / 4);
				for (size_t t = 0; t < m.size(); ++t) {
					const float d2 = c[4 * t + 2], d3 = c[4 * t + 3];
					m[t] = 1000.0f - (d2 - 30.0f) * (d2 - 30.0f) - 2.0f * (d3 + 8.0f) * (d3 + 8.0f);     /* maximum at d2 = 30, d3 = -8 */
				}
				if (ascans) ascans->assign(m.size() * lines * 32, 1.5f);
				return m;
			}
		};
		int calls = 0; std::vector<unsigned> linesSeen; const void* rawSeen = nullptr;
		DispersionEstimationEngine<FakeSweep> eng(FakeSweep{&calls, &linesSeen, &rawSeen}, 0.0f, 97.0f);
		DispersionEstimatorParameters prm;
		prm.numberOfCenterAscans = 10; prm.numberOfDispersionSamples = 100; prm.d2start = -100; prm.d2end = 100; prm.d3start = -20; prm.d3end = 20; prm.autoCalcD1 = true;
		eng.setParams(prm);
		std::vector<unsigned short> frame(64 * 50, 7);
		eng.startDispersionEstimation(frame.data(), 12, 64, 50);
		CHECK(calls == 3 && linesSeen[0] == 10 && linesSeen[1] == 10 && linesSeen[2] == 1);                      /* d2 sweep, d3 sweep, the two plotted A-scans */
		CHECK(rawSeen == reinterpret_cast<const char*>(frame.data()) + (size_t)20 * 64 * 2);                     /* center block: offset (50 - 10) / 2 lines */
		CHECK(std::fabs(eng.bestD2 - 30.0) < 1e-9 && std::fabs(eng.bestD3 - (-8.0)) < 1e-9);                     /* both on the trial grids (step 2 and 0.4) */
		CHECK(eng.metricsD2.size() == 100 && eng.metricsD3.size() == 100 && eng.metricsD2[0].first == -100.0);
		CHECK(std::fabs(eng.metricsD2[99].first - 98.0) < 1e-9);                                                  /* start + 99 steps: the end value itself is never tried */
		CHECK(std::fabs(eng.calculatedD1 - (-22.0)) < 1e-9);
		CHECK(eng.ascanWithBestDispersion.size() == 32 && eng.ascanWithoutDispersionCompensation.size() == 32);
		/* quirk of the reference kept: "best" starts at a metric of 0 with a strict '<' -- a metric that is never positive selects nothing */
		struct NegSweep {
			std::vector<float> operator()(const void*, unsigned lines, const std::vector<float>& c, std::vector<float>* a) {
				if (a) a->assign(c.size() / 4 * lines * 32, 0.0f);
				return std::vector<float>(c.size() / 4, -1.0f);
			}
		};
		DispersionEstimationEngine<NegSweep> neg(NegSweep{}, 0.0f, 0.0f);
		neg.setParams(prm);
		neg.startDispersionEstimation(frame.data(), 12, 64, 50);
		CHECK(neg.bestD2 == 0.0 && neg.bestD3 == 0.0 && neg.bestMetricValueD2 == 0.0);
		/* fewer lines than center A-scans: the whole frame, no offset (:37-47) */
		calls = 0; linesSeen.clear();
		eng.startDispersionEstimation(frame.data(), 12, 64, 4);
		CHECK(linesSeen[0] == 4 && rawSeen == frame.data());
	}
	/* the estimator path's own window (processor.tpp:126-133 with T = float) */
	{
		std::vector<float> w = cpuPathWindow(1024);
		CHECK(w.size() == 1024 && w[0] == 0.0f && std::fabs(w[1023]) < 1e-6f && std::fabs(w[511] - 1.0f) < 1e-5f);
		if (argc > 2) { FILE* f = std::fopen(argv[2], "wb"); CHECK(f); std::fwrite(w.data(), 4, w.size(), f); std::fclose(f); }
	}
	/* settings INI (QSettings vocabulary of the reference): argv[3] = a settings file, fields echoed for tests/test_host_mirror.py */
	if (argc > 3) {
		OctAlgorithmParameters q;
		AcquisitionParams acq;
		OctAlgorithmParameters::VirtualOctSettings vs;
		CHECK(OctAlgorithmParameters::fromIni(argv[3], &q, &acq, &vs));
		CHECK(!OctAlgorithmParameters::fromIni(std::string(argv[3]) + ".missing", &q));
		std::printf("INI {\"bitshift\": %d, \"bscanFlip\": %d, \"signalLogScaling\": %d, \"sinusoidalScanCorrection\": %d, \"signalGrayscaleMin\": %.9g, "
		            "\"signalGrayscaleMax\": %.9g, \"signalMultiplicator\": %.9g, \"signalAddend\": %.9g, \"backgroundRemoval\": %d, "
		            "\"rollingAverageWindowSize\": %d, \"resampling\": %d, \"resamplingInterpolation\": %d, \"c\": [%.9g, %.9g, %.9g, %.9g], "
		            "\"dispersionCompensation\": %d, \"d\": [%.9g, %.9g, %.9g, %.9g], \"windowing\": %d, \"window\": %d, \"windowFillFactor\": %.9g, "
		            "\"windowCenter\": %.9g, \"fixedPatternNoiseRemoval\": %d, \"continuousFixedPatternNoiseDetermination\": %d, "
		            "\"bscansForNoiseDetermination\": %u, \"postProcessBackgroundRemoval\": %d, \"postProcessBackgroundWeight\": %.9g, "
		            "\"postProcessBackgroundOffset\": %.9g, \"streamToHost\": %d, \"streamingBuffersToSkip\": %u, \"bitDepth\": %u, \"samplesPerLine\": %u, "
		            "\"ascansPerBscan\": %u, \"bscansPerBuffer\": %u, \"buffersPerVolume\": %u, \"buffersFromFile\": %d, \"bscanOffset\": %u, \"syncWithProcessing\": %d}\n",
		            q.p.bitshift, q.p.bscanFlip, q.p.signalLogScaling, q.p.sinusoidalScanCorrection, q.p.signalGrayscaleMin, q.p.signalGrayscaleMax,
		            q.p.signalMultiplicator, q.p.signalAddend, q.p.backgroundRemoval, q.p.rollingAverageWindowSize, q.p.resampling, q.p.resamplingInterpolation,
		            q.c[0], q.c[1], q.c[2], q.c[3], q.p.dispersionCompensation, q.d[0], q.d[1], q.d[2], q.d[3], q.p.windowing, q.window, q.windowFillFactor,
		            q.windowCenter, q.p.fixedPatternNoiseRemoval, q.p.continuousFixedPatternNoiseDetermination, q.p.bscansForNoiseDetermination,
		            q.p.postProcessBackgroundRemoval, q.p.postProcessBackgroundWeight, q.p.postProcessBackgroundOffset, q.p.streamToHost, q.p.streamingBuffersToSkip,
		            acq.bitDepth, acq.samplesPerLine, acq.ascansPerBscan, acq.bscansPerBuffer, acq.buffersPerVolume, vs.buffersFromFile, vs.bscanOffset, (int)vs.syncWithProcessing);
	}
	std::puts("host mirror ok");
	return 0;
}
