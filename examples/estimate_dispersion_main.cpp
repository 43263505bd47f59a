/* Dispersion estimation on the GPU, host side in C++ (include/octb200_host.hpp): the search of the reference's Dispersion Estimator
 * extension (dispersionestimationengine.cpp:21-116) with every sweep as one octb200_dispersion_sweep call.
 *   ./estimate <raw frame file> <samplesPerLine> <linesPerFrame> <bitDepth> <d2start> <d2end> <d3start> <d3end> <samples> <centerAscans>
 * Processing settings are the published benchmark settings; prints one JSON line with the best d2 / d3 and their metric values. */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "octb200_host.hpp"

using namespace octb200::host;

int main(int argc, char** argv) {
	if (argc < 11) { std::fprintf(stderr, "usage: %s file N lines bits d2start d2end d3start d3end samples centerAscans\n", argv[0]); return 2; }
	const unsigned n = (unsigned)std::atoi(argv[2]), lines = (unsigned)std::atoi(argv[3]), bits = (unsigned)std::atoi(argv[4]);
	std::vector<unsigned short> frame((size_t)n * lines);
	FILE* f = std::fopen(argv[1], "rb");
	if (!f || std::fread(frame.data(), 2, frame.size(), f) != frame.size()) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
	std::fclose(f);

	OctAlgorithmParameters q = OctAlgorithmParameters::benchmark(n);
	AcquisitionParams acq; acq.samplesPerLine = n; acq.ascansPerBscan = lines; acq.bscansPerBuffer = 1; acq.buffersPerVolume = 1; acq.bitDepth = bits;
	OctPipeline pipe(OCTB200_FFT_FUSED);
	if (!pipe.initializeCuda(nullptr, nullptr, acq, &q)) { std::fprintf(stderr, "initializeCuda failed: %s\n", pipe.lastError().c_str()); return 3; }

	DispersionEstimatorParameters prm;
	prm.d2start = std::atof(argv[5]); prm.d2end = std::atof(argv[6]); prm.d3start = std::atof(argv[7]); prm.d3end = std::atof(argv[8]);
	prm.numberOfDispersionSamples = std::atoi(argv[9]); prm.numberOfCenterAscans = std::atoi(argv[10]);
	prm.useLinearAscans = true; prm.numberOfAscanSamplesToIgnore = 15; prm.sharpnessMetric = OCTB200_METRIC_PEAK_VALUE; prm.autoCalcD1 = true;
	try {
		PipelineSweep sweep(pipe, q, n, prm);
		DispersionEstimationEngine<std::reference_wrapper<PipelineSweep>> eng(std::ref(sweep), q.d[0], q.d[1]);
		eng.setParams(prm);
		eng.startDispersionEstimation(frame.data(), bits, n, lines);
		std::printf("{\"bestD2\": %.9g, \"bestD3\": %.9g, \"bestMetricValueD2\": %.9g, \"bestMetricValueD3\": %.9g, \"calculatedD1\": %.9g, \"launches\": %llu}\n",
		            eng.bestD2, eng.bestD3, eng.bestMetricValueD2, eng.bestMetricValueD3, eng.calculatedD1, pipe.launchCount());
	} catch (const std::exception& e) {
		std::fprintf(stderr, "estimation failed: %s\n", e.what());
		return 3;
	}
	pipe.cleanupCuda();
	return 0;
}
