/*
 * adapter_harness.cpp -- the SAME headless driver API as ref_cuda.cpp (refcuda_*), but linked against
 * integration/octproz_kernels_adapter.cpp + liboctb200.so instead of the reference's cuda_code.cu.
 * It calls initializeCuda / octCudaPipeline / cleanupCuda by the reference's own names with the reference's own
 * OctAlgorithmParameters singleton (octalgorithmparameters.cpp compiled in place): the drop-in check.
 * TEST INFRASTRUCTURE ONLY.
 */
#define REFCUDA_NO_MEANLINE_GLOBAL 1
#include "ref_cuda.cpp"
