/* k_fused_cplx.cu -- own FFT + fused epilogue reading float2 from HBM (OCTB200_FFT_SPLIT), and the dispatcher */
#include "k_fused_launch.cuh"
namespace octb200 {
cudaError_t launch_fused_raw_r1(int sa, bool roll, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw_r2(int sa, bool roll, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw12(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw_r1_conv(int sa, bool roll, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw_r2_conv(int sa, bool roll, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw12_conv(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw8(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st);
cudaError_t launch_fused_raw32(int R, int sa, const FusedArgs& a, int smCount, cudaStream_t st);

cudaError_t launch_fused(int R, int sa, bool roll, int src, const FusedArgs& a, int smCount, cudaStream_t st) {
	if (src == SRC_CPLX) {
		if (a.convOut) return cudaErrorInvalidValue;      /* the converted output is only folded into the raw-source kernels */
		if (R == 1) return launch_fused_t<1, SA_NONE, false, SRC_CPLX>(a, smCount, st);
		return launch_fused_t<2, SA_NONE, false, SRC_CPLX>(a, smCount, st);
	}
	if (src == SRC_RAW8 || src == SRC_RAW32) {
		if (a.convOut || roll) return cudaErrorInvalidConfiguration;      /* converted output / rolling mean: u16 kernels only */
		return src == SRC_RAW8 ? launch_fused_raw8(R, sa, a, smCount, st) : launch_fused_raw32(R, sa, a, smCount, st);
	}
	if (a.convOut) {
		if (src == SRC_RAW12P) return roll ? cudaErrorInvalidConfiguration : launch_fused_raw12_conv(R, sa, a, smCount, st);
		if (R == 1) return launch_fused_raw_r1_conv(sa, roll, a, smCount, st);
		return launch_fused_raw_r2_conv(sa, roll, a, smCount, st);
	}
	if (src == SRC_RAW12P) return roll ? cudaErrorInvalidConfiguration : launch_fused_raw12(R, sa, a, smCount, st);
	if (R == 1) return launch_fused_raw_r1(sa, roll, a, smCount, st);
	return launch_fused_raw_r2(sa, roll, a, smCount, st);
}

void fused_launch_shape(int R, int sa, bool roll, int src, int HB, int HA, int smCount, int lines, int* grid, int* threads, int* smem) {
	if (src == SRC_CPLX) { sa = SA_NONE; roll = false; }
	if (sa == SA_LINEAR) sa = SA_CUBIC;
	const int groups = fused_pick_groups(R, sa, roll, src, HB, HA);
	const FusedSmem L = fused_smem_layout(R, sa, roll, src, HB, HA, groups > 0 ? groups : 1);
	int g = smCount; const int mg = groups > 0 ? (lines + groups - 1) / groups : 1;
	if (g > mg) g = mg; if (g < 1) g = 1;
	if (grid) *grid = g; if (threads) *threads = groups * R * 32; if (smem) *smem = L.total;
}
}
