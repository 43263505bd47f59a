/* k_fused_p12.cu -- instantiations of the fused kernel for 12-bit packed input (SRC_RAW12P): 4-tap and plain stages, no rolling mean.
 * (Lanczos halos and the rolling-mean prefix pass read the raw slot as u16: for those the host unpacks first, octb200.cu.) */
#include "k_fused_launch.cuh"
namespace octb200 {
cudaError_t launch_fused_raw12(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st) {
	if (sa == SA_CUBIC || sa == SA_LINEAR)
		return R == 1 ? launch_fused_t<1, SA_CUBIC, false, SRC_RAW12P>(a, smCount, st) : launch_fused_t<2, SA_CUBIC, false, SRC_RAW12P>(a, smCount, st);
	if (sa == SA_NONE)
		return R == 1 ? launch_fused_t<1, SA_NONE, false, SRC_RAW12P>(a, smCount, st) : launch_fused_t<2, SA_NONE, false, SRC_RAW12P>(a, smCount, st);
	return cudaErrorInvalidConfiguration;
}
}
