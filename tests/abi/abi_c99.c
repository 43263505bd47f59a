/* The C ABI used from plain C99 (no C++, no CUDA headers): what a cgo / JNI / FFI binding of include/octb200.h sees.
 * Built and run by tests/test_abi.py; needs no GPU (create must fail loudly without one or on bad geometry). */
#include <stdio.h>
#include <string.h>

#include "octb200.h"

int main(void) {
	octb200_params prm;
	octb200_config cfg;
	octb200_pipeline* h = NULL;
	int rc;

	printf("version %d sizeof_config %u sizeof_params %u\n", octb200_version(), (unsigned)sizeof(octb200_config), (unsigned)sizeof(octb200_params));
	octb200_default_params(&prm);
	if (prm.signalGrayscaleMax != 60.0f || prm.signalMultiplicator != 1.0f || prm.bscansForNoiseDetermination != 1u) return 2;

	memset(&cfg, 0, sizeof(cfg));
	cfg.samplesPerLine = 7; cfg.ascansPerBscan = 4; cfg.bscansPerBuffer = 1; cfg.buffersPerVolume = 1; cfg.bitDepth = 12; cfg.device = -1;
	rc = octb200_create(&cfg, &h);
	if (rc != OCTB200_ERR_INVALID || h != NULL) return 3;
	if (strstr(octb200_last_error(NULL), "geometry") == NULL) return 4;
	if (octb200_create(NULL, &h) != OCTB200_ERR_INVALID) return 5;
	/* every entry point taking a handle rejects NULL instead of crashing */
	if (octb200_process_host(NULL, NULL) != OCTB200_ERR_INVALID) return 6;
	if (octb200_set_params(NULL, &prm) != OCTB200_ERR_INVALID) return 7;
	if (octb200_sync(NULL) != OCTB200_ERR_INVALID) return 8;
	if (octb200_destroy(NULL) != OCTB200_OK && octb200_destroy(NULL) != OCTB200_ERR_INVALID) return 9;
	printf("flags %d\n", (int)OCTB200_FLAG_SEPARATE_CONVERSION);
	puts("abi ok");
	return 0;
}
