/* k_fused_c8c32.cu -- instantiations of the fused kernel for u8 and u32 containers (SRC_RAW8 / SRC_RAW32): 4-tap and plain stages, no
 * rolling mean.  (Lanczos halos and the rolling-mean prefix pass read the raw slot as u16: those configurations take the split chain.) */
#include "k_fused_launch.cuh"
namespace octb200 {
template <int SRC>
static cudaError_t launch_container(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st) {
	if (sa == SA_CUBIC || sa == SA_LINEAR)
		return R == 1 ? launch_fused_t<1, SA_CUBIC, false, SRC>(a, smCount, st) : launch_fused_t<2, SA_CUBIC, false, SRC>(a, smCount, st);
	if (sa == SA_NONE)
		return R == 1 ? launch_fused_t<1, SA_NONE, false, SRC>(a, smCount, st) : launch_fused_t<2, SA_NONE, false, SRC>(a, smCount, st);
	return cudaErrorInvalidConfiguration;
}
cudaError_t launch_fused_raw8(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st) { return launch_container<SRC_RAW8>(R, sa, a, smCount, st); }
cudaError_t launch_fused_raw32(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st) { return launch_container<SRC_RAW32>(R, sa, a, smCount, st); }
}
