/* k_fused_launch.cuh -- per-instantiation launcher, included by the k_fused_*.cu translation units */
#pragma once
#include "k_aux.cuh"

namespace octb200 {

inline int fused_pick_groups(int R, int sa, bool roll, int src, int HB, int HA) {
	const int maxThreads = fused_max_threads(R, sa);
	int groups = maxThreads / 32 / R;
	while (groups > 0 && fused_smem_layout(R, sa, roll, src, HB, HA, groups).total > 227 * 1024) --groups;
	if (R == 2 && groups > 15) groups = 15;     /* named barriers 1..15 */
	return groups;
}

template <int R, int SA, bool ROLL, int SRC, bool CONV = false>
cudaError_t launch_fused_t(const FusedArgs& a, int smCount, cudaStream_t st) {
	if (CONV != (a.convOut != nullptr)) return cudaErrorInvalidValue;
	const int groups = fused_pick_groups(R, SA, ROLL, SRC, a.HB, a.HA);
	if (groups < 1) return cudaErrorInvalidConfiguration;
	const FusedSmem L = fused_smem_layout(R, SA, ROLL, SRC, a.HB, a.HA, groups);
	auto k = oct_fused_kernel<R, SA, ROLL, SRC, CONV>;
	cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
	if (e != cudaSuccess) return e;
	int grid = smCount;
	const int maxGrid = (a.lines + groups - 1) / groups;
	if (grid > maxGrid) grid = maxGrid;
	if (grid < 1) grid = 1;
	if (a.pdl) {
		/* programmatic dependent launch: this grid's prologue may overlap the tail of the previous main launch (oct_device.cuh) */
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(grid, a.trials > 1 ? a.trials : 1); cfg.blockDim = dim3(groups * R * 32); cfg.dynamicSmemBytes = (size_t)L.total; cfg.stream = st;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		return cudaLaunchKernelEx(&cfg, k, a);
	}
	k<<<dim3(grid, a.trials > 1 ? a.trials : 1), groups * R * 32, L.total, st>>>(a);
	return cudaGetLastError();
}

}  // namespace octb200
