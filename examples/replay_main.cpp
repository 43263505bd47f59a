/* Headless Virtual-OCT-System replay through the B200 pipeline, host side in C++ (include/octb200_host.hpp over include/octb200.h):
 * the reference's acquisition thread + processing loop without Qt.
 *   g++ -std=c++17 -O2 -pthread -Iinclude examples/replay_main.cpp -Loctproz_b200 -loctb200 -Wl,-rpath,$PWD/octproz_b200 -o replay
 *   ./replay <raw file> <samplesPerLine> <ascansPerBscan> <bscansPerBuffer> <bitDepth> <buffers>      (published benchmark settings)
 *   ./replay --ini <settings.ini of the reference> <raw file> <buffers>                              (geometry + processing from the file)  Prints one JSON line: the sidebar statistics of processing.cpp:194-207,
 * the kernel launches per buffer and a checksum of the last processed buffer. */
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "octb200_host.hpp"

using namespace octb200::host;

int main(int argc, char** argv) {
	/* second form: everything from the reference's own settings file --  ./replay --ini settings.ini <raw file> <buffers>  */
	const bool fromIni = argc >= 5 && std::string(argv[1]) == "--ini";
	if (!fromIni && argc < 7) {
		std::fprintf(stderr, "usage: %s file samplesPerLine ascansPerBscan bscansPerBuffer bitDepth buffers\n       %s --ini settings.ini file buffers\n", argv[0], argv[0]);
		return 2;
	}
	unsigned n, a, b, bits;
	long long buffers;
	OctAlgorithmParameters q;
	OctAlgorithmParameters::VirtualOctSettings vs;
	const char* file;
	if (fromIni) {
		AcquisitionParams acq;
		if (!OctAlgorithmParameters::fromIni(argv[2], &q, &acq, &vs)) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
		n = acq.samplesPerLine; a = acq.ascansPerBscan; b = acq.bscansPerBuffer; bits = acq.bitDepth;
		if (!q.updateCurves(n)) { std::fprintf(stderr, "bad curve parameters in %s\n", argv[2]); return 2; }
		file = argv[3]; buffers = std::atoll(argv[4]);
	} else {
		n = (unsigned)std::atoi(argv[2]); a = (unsigned)std::atoi(argv[3]); b = (unsigned)std::atoi(argv[4]); bits = (unsigned)std::atoi(argv[5]);
		buffers = std::atoll(argv[6]);
		q = OctAlgorithmParameters::benchmark(n);
		file = argv[1];
	}
	VirtualOCTSystem vos(file, bits, n, a, b, 1);
	vos.buffersFromFile = vs.buffersFromFile; vos.bscanOffset = vs.bscanOffset; vos.waitTimeUs = vs.waitTimeUs; vos.syncWithProcessing = vs.syncWithProcessing;
	q.p.streamToHost = 0;          /* no consumer registered for the converted buffers in this program */
	OctPipeline pipe(OCTB200_FFT_AUTO);
	bool ok = false;
	ProcessingStats st;
	try {
		st = replay(vos, pipe, q, buffers, &ok);
	} catch (const std::exception& e) {
		std::fprintf(stderr, "replay failed: %s\n", e.what());
		return 3;
	}
	if (!ok) { std::fprintf(stderr, "initializeCuda failed: %s\n", pipe.lastError().c_str()); return 3; }
	std::vector<float> out((size_t)(n / 2) * a * b);
	pipe.copyOutput(out.data(), 0);
	double sum = 0.0;
	for (float v : out) sum += (double)v;
	std::printf("{\"processed_buffers\": %lld, \"buffers_per_s\": %.3f, \"ascans_per_s\": %.1f, \"MB_per_s\": %.1f, \"launches\": %llu, \"output_sum\": %.6f}\n",
	            st.processedBuffers, st.buffersPerSecond, st.ascansPerSecond, st.dataThroughputMBs, pipe.launchCount(), sum);
	pipe.cleanupCuda();
	return 0;
}
