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
	std::puts("host mirror ok");
	return 0;
}
