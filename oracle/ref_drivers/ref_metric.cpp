/*
 * ref_metric.cpp -- C entry point around the reference's OWN A-scan sharpness metric,
 * AscanMetricCalculator::calculateMetric (octproz-dispersion-estimator-extension/src/ascanmetriccalculator.cpp:22-128),
 * compiled in place from /root/reference against the Qt stand-ins of oracle/shim.
 * TEST INFRASTRUCTURE ONLY: pins oracle.ascan_metric and the GPU metric kernel.
 */
#include "ascanmetriccalculator.h"

extern "C" float refmetric_calculate(const float* data, int total, int samplesPerLine, int metric, double threshold, int ignore) {
	DispersionEstimatorParameters prm{};
	prm.sharpnessMetric = (ASCAN_SHARPNESS_METRIC)metric;
	prm.metricThreshold = threshold;
	prm.numberOfAscanSamplesToIgnore = ignore;
	AscanMetricCalculator calc(prm);
	QVector<float> v(data, data + total);
	return calc.calculateMetric(v, samplesPerLine);
}
