/* k_fused_p12.cu -- instantiations of the fused kernel for 12-bit packed input (SRC_RAW12P): 4-tap and plain stages, no rolling mean.
 * (Lanczos halos and the rolling-mean prefix pass read the raw slot as u16: for those the host unpacks first, octb200.cu.) */
#include "k_fused_launch.cuh"
/* compiled twice (Makefile): OCT_FUSED_CONV = 0 -> the plain kernels, 1 -> the same kernels with floatToOutput folded into the epilogue */
#ifndef OCT_FUSED_CONV
#define OCT_FUSED_CONV 0
#endif
namespace octb200 {
constexpr bool CV = OCT_FUSED_CONV != 0;
#if OCT_FUSED_CONV
#define launch_fused_raw12 launch_fused_raw12_conv
#endif
cudaError_t launch_fused_raw12(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st) {
	if (sa == SA_CUBIC || sa == SA_LINEAR)
		return R == 1 ? launch_fused_t<1, SA_CUBIC, false, SRC_RAW12P, CV>(a, smCount, st) : launch_fused_t<2, SA_CUBIC, false, SRC_RAW12P, CV>(a, smCount, st);
	if (sa == SA_NONE)
		return R == 1 ? launch_fused_t<1, SA_NONE, false, SRC_RAW12P, CV>(a, smCount, st) : launch_fused_t<2, SA_NONE, false, SRC_RAW12P, CV>(a, smCount, st);
	return cudaErrorInvalidConfiguration;
}
}
