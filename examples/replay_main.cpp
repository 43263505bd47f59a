/* Headless Virtual-OCT-System replay through the B200 pipeline, host side in C++ (include/octb200_host.hpp over include/octb200.h):
 * the reference's acquisition thread + processing loop without Qt.
 *   g++ -std=c++17 -O2 -pthread -Iinclude examples/replay_main.cpp -Loctproz_b200 -loctb200 -Wl,-rpath,$PWD/octproz_b200 -o replay
 *   ./replay <raw file> <samplesPerLine> <ascansPerBscan> <bscansPerBuffer> <bitDepth> <buffers>
 * Processing parameters are the published benchmark settings.  Prints one JSON line: the sidebar statistics of processing.cpp:194-207,
 * the kernel launches per buffer and a checksum of the last processed buffer. */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "octb200_host.hpp"

using namespace octb200::host;

int main(int argc, char** argv) {
	if (argc < 7) { std::fprintf(stderr, "usage: %s file samplesPerLine ascansPerBscan bscansPerBuffer bitDepth buffers\n", argv[0]); return 2; }
	const unsigned n = (unsigned)std::atoi(argv[2]), a = (unsigned)std::atoi(argv[3]), b = (unsigned)std::atoi(argv[4]), bits = (unsigned)std::atoi(argv[5]);
	const long long buffers = std::atoll(argv[6]);
	VirtualOCTSystem vos(argv[1], bits, n, a, b, 1);
	OctAlgorithmParameters q = OctAlgorithmParameters::benchmark(n);
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
