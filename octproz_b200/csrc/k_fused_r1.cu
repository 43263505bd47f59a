/* k_fused_r1.cu -- instantiations of the fused kernel for N = 1024 (raw u16 source) */
#include "k_fused_launch.cuh"
/* compiled twice (Makefile): OCT_FUSED_CONV = 0 -> the plain kernels, 1 -> the same kernels with floatToOutput folded into the epilogue */
#ifndef OCT_FUSED_CONV
#define OCT_FUSED_CONV 0
#endif
namespace octb200 {
constexpr bool CV = OCT_FUSED_CONV != 0;
#if OCT_FUSED_CONV
#define launch_fused_raw_r1 launch_fused_raw_r1_conv
#endif
cudaError_t launch_fused_raw_r1(int sa, bool roll, const FusedArgs& a, int smCount, cudaStream_t st) {
	if (sa == SA_CUBIC || sa == SA_LINEAR) return   /* linear = the same 4-tap kernel with weights (0, 1-t, t, 0) */ roll ? launch_fused_t<1, SA_CUBIC, true, SRC_RAW16, CV>(a, smCount, st) : launch_fused_t<1, SA_CUBIC, false, SRC_RAW16, CV>(a, smCount, st);
	if (sa == SA_NONE) return roll ? launch_fused_t<1, SA_NONE, true, SRC_RAW16, CV>(a, smCount, st) : launch_fused_t<1, SA_NONE, false, SRC_RAW16, CV>(a, smCount, st);
	return roll ? launch_fused_t<1, SA_LANCZOS, true, SRC_RAW16, CV>(a, smCount, st) : launch_fused_t<1, SA_LANCZOS, false, SRC_RAW16, CV>(a, smCount, st);
}
}
