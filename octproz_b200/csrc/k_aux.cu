/*
 * k_aux.cu -- the lighter kernels around the fused hot kernel (all sm_100a, hand written):
 *   fill_phase          : fillDispersivePhase                           (cuda_code.cu:624-634)
 *   oct_pre_kernel      : fused pre-FFT kernel, any N / container       (cuda_code.cu:109-489) -> float2 FFT input
 *   oct_post_kernel     : FPN subtract + truncate + log/lin + flip      (cuda_code.cu:567-584,699-807) after cuFFT
 *   fpn_minvar_kernel   : getMinimumVarianceMean                        (cuda_code.cu:523-565)
 *   sinusoidal_kernel   : sinusoidalScanCorrection (+ folded background)(cuda_code.cu:491-514,757-767)
 *   ppbg_*              : getPostProcessBackground / ...Removal         (cuda_code.cu:743-767)
 *   bscan/enface/volume : updateDisplayed*                              (cuda_code.cu:810-941)
 *   float_to_output     : floatToOutput                                 (cuda_code.cu:943-967)
 */
#include "k_aux.cuh"
#include <cfloat>

namespace octb200 {

/* ------------------------------------------------------------------ dispersion phasor */
__global__ void fill_phase_kernel(float2* __restrict__ ph, const float* __restrict__ phase, int n) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		/* cosf(1.0*phase) / sinf(..)*1 under --use_fast_math == __cosf/__sinf (cuda_code.cu:631-632, cuda.pri:54) */
		const float a = phase[i];
		ph[i] = make_float2(__cosf(a), __sinf(a));
	}
}
void launch_fill_phase(float2* ph, const float* phase, int n, cudaStream_t st) {
	fill_phase_kernel<<<(n + 127) / 128, 128, 0, st>>>(ph, phase, n);
}

/* ------------------------------------------------------------------ generic pre-FFT kernel */
template <typename RawT> struct RawTraits;
template <> struct RawTraits<uint8_t> { static constexpr int kBytes = 1; };
template <> struct RawTraits<uint16_t> { static constexpr int kBytes = 2; };
template <> struct RawTraits<uint32_t> { static constexpr int kBytes = 4; };

template <typename RawT>
__device__ __forceinline__ float convert_raw(RawT v, int shiftBits) {
	if constexpr (sizeof(RawT) == 4) {
		if (shiftBits) return (float)((double)v / 4294967296.0);   /* cuda_code.cu:144 */
		return __uint2float_rd(v);                                  /* cuda_code.cu:124 */
	} else {
		return __uint2float_rd((unsigned)v >> shiftBits);           /* cuda_code.cu:118-121,138-141 */
	}
}

/* per-warp shared layout of the pre kernel */
__host__ __device__ inline int pre_warp_bytes(int SE, int rawBytes, bool roll) {
	return 2 * align_up(SE * rawBytes, 128) + align_up((FSLOT_PAD + SE) * 4, 128) + (roll ? align_up((SE + 1) * 8, 128) : 0) + 128;
}

template <typename RawT, int SA, bool ROLL>
__global__ void __launch_bounds__(256) oct_pre_kernel(const PreArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int RB = sizeof(RawT);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int N = a.N, SE = a.HB + N + a.HA;
	unsigned char* wbase = smem + warp * pre_warp_bytes(SE, RB, ROLL);
	const int slotBytes = align_up(SE * RB, 128);
	unsigned char* slots[2] = { wbase, wbase + slotBytes };
	float* fslot = reinterpret_cast<float*>(wbase + 2 * slotBytes) + FSLOT_PAD;
	unsigned long long* prefix = reinterpret_cast<unsigned long long*>(wbase + 2 * slotBytes + align_up((FSLOT_PAD + SE) * 4, 128));
	uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + pre_warp_bytes(SE, RB, ROLL) - 128);

	const int G = gridDim.x * (blockDim.x >> 5);
	const int g0 = blockIdx.x * (blockDim.x >> 5) + warp;
	const RawT* raw = reinterpret_cast<const RawT*>(a.raw);

	auto issue = [&](int gline, int s) {
		const long long lo = (long long)gline * N - a.HB, hi = (long long)gline * N + N + a.HA;
		const long long clo = lo < 0 ? 0 : lo, chi = hi > a.totalSamples ? a.totalSamples : hi;
		RawT* dst = reinterpret_cast<RawT*>(slots[s]);
		if (a.useBulk) {
			if (lane == 0) {
				for (long long q = 0; q < clo - lo; ++q) dst[q] = 0;
				for (long long q = chi - lo; q < hi - lo; ++q) dst[q] = 0;
				const uint32_t bytes = (uint32_t)((chi - clo) * RB);
				mbar_arrive_expect_tx(&bars[s], bytes);
				bulk_g2s(dst + (clo - lo), raw + clo, bytes, &bars[s]);
			}
		} else {
			/* unaligned geometry: plain coalesced loads straight into the slot */
			for (long long q = lo + lane; q < hi; q += 32) dst[q - lo] = (q >= 0 && q < a.totalSamples) ? raw[q] : (RawT)0;
		}
	};

	if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
	__syncwarp();
	if (g0 < a.lines) issue(g0, 0);
	if (a.useBulk && g0 + G < a.lines) issue(g0 + G, 1);

	int it = 0;
	for (int gline = g0; gline < a.lines; gline += G, ++it) {
		const int s = a.useBulk ? (it & 1) : 0;
		if (a.useBulk) mbar_wait(&bars[s], (uint32_t)((it >> 1) & 1));
		else __syncwarp();
		const RawT* rs = reinterpret_cast<const RawT*>(slots[s]);
		if constexpr (ROLL) {
			unsigned long long carry = 0;
			if (lane == 0) prefix[0] = 0;
			for (int c = 0; c < SE; c += 32) {
				const int q = c + lane;
				unsigned long long x = 0;
				if (q < SE) x = (sizeof(RawT) == 4) ? (unsigned long long)rs[q] : (unsigned long long)((unsigned)rs[q] >> a.shiftBits);
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
				if (q < SE) prefix[q + 1] = carry + x;
				carry += __shfl_sync(0xffffffffu, x, 31);
			}
		}
		for (int q = lane; q < SE; q += 32) fslot[q] = convert_raw<RawT>(rs[q], a.shiftBits);
		__syncwarp();
		if (a.useBulk) { if (gline + 2 * G < a.lines) issue(gline + 2 * G, s); }
		if constexpr (ROLL) {
			const int W = a.W;
			for (int q = lane; q < SE; q += 32) {
				int lo, hi;
				if (q < a.HB) { lo = 0; hi = a.HB - 1; }
				else if (q >= a.HB + N) { lo = a.HB + N; hi = SE - 1; }
				else { lo = a.HB; hi = a.HB + N - 1; }
				const int ss = max(lo, q - W + 1), e = min(hi, q + W);
				const unsigned long long d = prefix[e + 1] - prefix[ss];
				float sum;
				if (sizeof(RawT) == 4 && a.shiftBits) sum = (float)((double)d / 4294967296.0);
				else sum = (float)d;
				fslot[q] -= __fdividef(sum, (float)(e - ss + 1));
			}
			__syncwarp();
		}
		if constexpr (SA == SA_CUBIC) {
			if (lane == 0) fslot[a.HB - 1] = fslot[a.HB + 1];      /* mirrored first tap of the cubic (cuda_code.cu:284) */
			__syncwarp();
		}
		const float* f = fslot + a.HB;
		const int shift = (SA == SA_LANCZOS && gline == 0) ? 8 : 0;
		float2* o = a.out + (size_t)gline * N;
		/* four table reads in flight per lane: the loop was bound by the latency of one L1/L2 read per sample (ncu: long_scoreboard 5.9
		 * warps per issue cycle, profiles/r02m_fused_ncu_summary.md) */
#pragma unroll 4
		for (int m = lane; m < N; m += 32) {
			const float4 B = __ldg(a.lutB + m);
			float2 val;
			if constexpr (SA == SA_CUBIC) val = sample_cubic(f, B);
			else if constexpr (SA == SA_LINEAR) val = sample_linear(f, B);
			else if constexpr (SA == SA_NONE) val = sample_none(f, m, B);
			else val = sample_lanczos(f, shift, B);
			o[m] = val;
		}
		__syncwarp();
		if (!a.useBulk) { if (gline + G < a.lines) issue(gline + G, 0); }
	}
}

template <typename RawT, int SA, bool ROLL>
static cudaError_t launch_pre_t(const PreArgs& a, int smCount, cudaStream_t st) {
	const int SE = a.HB + a.N + a.HA;
	const int perWarp = pre_warp_bytes(SE, (int)sizeof(RawT), ROLL);
	int warps = 8;
	while (warps > 1 && warps * perWarp > 200 * 1024) warps >>= 1;
	if (warps * perWarp > 227 * 1024) return cudaErrorInvalidConfiguration;
	const int smem = warps * perWarp;
	auto k = oct_pre_kernel<RawT, SA, ROLL>;
	cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (e != cudaSuccess) return e;
	int ctasPerSm = (220 * 1024) / (smem + 1024);
	if (ctasPerSm < 1) ctasPerSm = 1;
	if (ctasPerSm > 4) ctasPerSm = 4;
	int grid = smCount * ctasPerSm;
	const int maxGrid = (a.lines + warps - 1) / warps;
	if (grid > maxGrid) grid = maxGrid;
	if (grid < 1) grid = 1;
	k<<<grid, warps * 32, smem, st>>>(a);
	return cudaGetLastError();
}

template <typename RawT>
static cudaError_t launch_pre_raw(const PreArgs& a, int sa, bool roll, int smCount, cudaStream_t st) {
	if (sa == SA_CUBIC) return roll ? launch_pre_t<RawT, SA_CUBIC, true>(a, smCount, st) : launch_pre_t<RawT, SA_CUBIC, false>(a, smCount, st);
	if (sa == SA_LINEAR) return roll ? launch_pre_t<RawT, SA_LINEAR, true>(a, smCount, st) : launch_pre_t<RawT, SA_LINEAR, false>(a, smCount, st);
	if (sa == SA_NONE) return roll ? launch_pre_t<RawT, SA_NONE, true>(a, smCount, st) : launch_pre_t<RawT, SA_NONE, false>(a, smCount, st);
	return roll ? launch_pre_t<RawT, SA_LANCZOS, true>(a, smCount, st) : launch_pre_t<RawT, SA_LANCZOS, false>(a, smCount, st);
}

cudaError_t launch_pre(const PreArgs& a, int rawBytes, int sa, bool roll, int smCount, cudaStream_t st) {
	if (rawBytes == 1) return launch_pre_raw<uint8_t>(a, sa, roll, smCount, st);
	if (rawBytes == 2) return launch_pre_raw<uint16_t>(a, sa, roll, smCount, st);
	return launch_pre_raw<uint32_t>(a, sa, roll, smCount, st);
}

/* ------------------------------------------------------------------ post kernel (after cuFFT) */
/* meanALineSubtraction + postProcessTruncateLog/Lin + bscanFlip (+ postProcessBackgroundRemoval) in one pass over the kept half of
 * the spectrum (cuda_code.cu:567-584, 699-741, 757-767, 787-807).  One warp per A-scan: line / B-scan / flip bookkeeping is
 * warp-uniform (no per-element 64-bit division), two bins per 16-byte load. */
__device__ __forceinline__ float post_one(float2 c, int z, const PostArgs& a) {
	if (a.epi.fpn) { const float2 m = __ldg(a.meanLine + z); c.x -= m.x; c.y -= m.y; }
	float o = scale_output(c.x, c.y, a.epi);
	if (a.epi.ppbg) o = saturate01(o - fmaf(a.epi.ppbgWeight, __ldg(a.ppbg + z), a.epi.ppbgOffset));
	return o;
}
__global__ void __launch_bounds__(256) oct_post_kernel(const PostArgs a, int vec) {
	const int H = a.N / 2;
	const int lane = threadIdx.x & 31, warpsPerBlock = blockDim.x >> 5;
	for (int line = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); line < a.lines; line += gridDim.x * warpsPerBlock) {
		int b = line / a.A, al = line - b * a.A;
		if (a.flip && (((unsigned)b + a.bscanBase) & 1u) == 0u && (unsigned)b + a.bscanBase < a.flipEnd) al = a.A - 1 - al;
		const float2* in = a.in + (size_t)line * a.N;
		float* out = a.out + ((size_t)b * a.A + al) * H;
		if (vec) {
#pragma unroll 4
			for (int z2 = lane; z2 < (H >> 1); z2 += 32) {
				const float4 c = *reinterpret_cast<const float4*>(in + 2 * z2);
				*reinterpret_cast<float2*>(out + 2 * z2) = make_float2(post_one(make_float2(c.x, c.y), 2 * z2, a), post_one(make_float2(c.z, c.w), 2 * z2 + 1, a));
			}
		} else {
			for (int z = lane; z < H; z += 32) out[z] = post_one(in[z], z, a);
		}
	}
}
cudaError_t launch_post(const PostArgs& a, int smCount, cudaStream_t st) {
	long long blocks = ((long long)a.lines + 7) / 8;
	if (blocks > (long long)smCount * 8) blocks = (long long)smCount * 8;
	if (blocks < 1) blocks = 1;
	/* 16-byte loads / 8-byte stores: even number of kept bins, line starts aligned */
	const int H = a.N / 2;
	const int vec = (H % 2 == 0) && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 7) == 0;
	oct_post_kernel<<<(int)blocks, 256, 0, st>>>(a, vec);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ fixed-pattern noise: minimum-variance mean */
/* block = 32 bins x 9 segments.  Each thread sums its segment in line order (same order as cuda_code.cu:544-551),
 * then the 9 candidates of a bin are compared with the reference's strict '<' starting from FLT_MAX. */
__global__ void __launch_bounds__(32 * 9) fpn_minvar_kernel(float2* __restrict__ meanLine, const float2* __restrict__ in,
                                                             int bins, int stride, int segW, float4* __restrict__ segStats) {
	__shared__ float sVar[9][32];
	__shared__ float2 sMean[9][32];
	const int bx = threadIdx.x, s = threadIdx.y;
	const int bin = blockIdx.x * 32 + bx;
	float var = 0.f; float2 mean = make_float2(0.f, 0.f);
	if (bin < bins && segW > 0) {
		/* every operation spelled out in the form nvcc --use_fast_math gives the reference's kernel (SASS of oracle/_ref/libref_cuda.so:
		 * FMUL dy dy -> FFMA dx dx -> FADD into sumXX; factor by MUFU.RCP; variance = FFMA(factor, sumXX, -FFMA(mx, mx, FMUL(my, my)))), so
		 * that the compiler's choice of contractions in THIS translation unit cannot move the ill-conditioned argmin */
		const float factor = __fdividef(1.0f, (float)segW);
		float sx = 0.f, sy = 0.f, sxx = 0.f;
		const float2* ptr = in + (size_t)s * segW * stride + bin;
		for (int j = 0; j < segW; ++j) {
			const float2 v = ptr[(size_t)j * stride];
			sx = __fadd_rn(sx, v.x); sy = __fadd_rn(sy, v.y);
			sxx = __fadd_rn(sxx, __fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y)));
		}
		mean.x = __fmul_rn(sx, factor); mean.y = __fmul_rn(sy, factor);
		var = __fmaf_rn(factor, sxx, -__fmaf_rn(mean.x, mean.x, __fmul_rn(mean.y, mean.y)));
		/* diagnostics (octb200_get_fpn_segment_stats): the nine candidates of every bin, { mean, variance, mean power } */
		if (segStats) segStats[(size_t)s * bins + bin] = make_float4(mean.x, mean.y, var, __fmul_rn(sxx, factor));
	}
	sVar[s][bx] = var; sMean[s][bx] = mean;
	__syncthreads();
	if (s == 0 && bin < bins) {
		float minVar = FLT_MAX; float2 best = make_float2(0.f, 0.f);
		if (segW > 0) {
#pragma unroll
			for (int i = 0; i < 9; ++i) if (sVar[i][bx] < minVar) { minVar = sVar[i][bx]; best = sMean[i][bx]; }
		}
		meanLine[bin] = best;
	}
}
cudaError_t launch_fpn_minvar(float2* meanLine, const float2* in, int bins, int stride, int height, float4* segStats, cudaStream_t st) {
	const int segW = height / 9;   /* FIXED_PATTERN_NOISE_REMOVAL_SEGMENTS, octalgorithmparameters.h:35; integer division cuda_code.cu:531 */
	fpn_minvar_kernel<<<(bins + 31) / 32, dim3(32, 9), 0, st>>>(meanLine, in, bins, stride, segW, segStats);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ sinusoidal scan correction */
/* cuda_code.cu:491-514 (+ the D2D copy of :1552, which disappears: the main kernel writes the source into a scratch slab).
 * out[l][k][j] = lerp(in[flat line l*A + n][j], in[flat line l*A + n + 1][j], x - n), x = curve[k], n = (int)x; the reference
 * addresses the slab flat (so n + 1 may be line 0 of the next B-scan) and leaves the very last line of the buffer untouched.
 * One warp per output line: x, n and the two source lines are warp-uniform, the line is moved as float4 (no per-element integer
 * division, which made the first version of this kernel instruction-bound: three 64-bit div/mod per output).  Neighbouring output
 * lines share their source lines (d curve / dk is 0.64 in the middle of the scan), consecutive warps take consecutive lines, so
 * the second read of a source line hits L1/L2: DRAM traffic is ~4 B read + 4 B written per output.
 * Optional in the same pass: post-process background removal (cuda_code.cu:757-767) and, CONV, the converted u16 line of
 * floatToOutput (cuda_code.cu:943-967; same round-toward-zero FMA as the fused kernel's epilogue, oct_tmem.cuh). */
__device__ __forceinline__ float sinus_lerp(float f0, float f1, float fr) { return f0 + (f1 - f0) * fr; }
__device__ __forceinline__ unsigned short to_u16_container(float v, float scale) {
	return (unsigned short)__float_as_uint(__fmaf_rz(__saturatef(v), scale, 8388608.0f));
}
template <bool CONV>
__global__ void __launch_bounds__(256) sinusoidal_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                         const float* __restrict__ curve, int H, int A, int lines, int vec,
                                                         int ppbgOn, const float* __restrict__ ppbg, float w, float o,
                                                         unsigned short* __restrict__ conv, float convScale) {
	const int lane = threadIdx.x & 31;
	const int warpsPerBlock = blockDim.x >> 5;
	for (int line = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); line < lines; line += gridDim.x * warpsPerBlock) {
		const int l = line / A, k = line - l * A;
		const bool last = line == lines - 1;      /* the last line keeps its uncorrected value (cuda_code.cu:499) */
		const float x = last ? 0.0f : __ldg(curve + k);
		const int n = (int)x;
		const float fr = x - (float)n;
		const float* s0 = last ? in + (size_t)line * H : in + ((size_t)l * A + n) * H;
		const float* s1 = last ? s0 : s0 + H;
		float* d = out + (size_t)line * H;
		unsigned short* c = CONV ? conv + (size_t)line * H : nullptr;
		if (vec) {
			const int H4 = H >> 2;
#pragma unroll 4
			for (int j4 = lane; j4 < H4; j4 += 32) {
				const float4 a = *reinterpret_cast<const float4*>(s0 + 4 * j4);
				const float4 b = *reinterpret_cast<const float4*>(s1 + 4 * j4);
				float4 r;
				if (last) r = a;
				else r = make_float4(sinus_lerp(a.x, b.x, fr), sinus_lerp(a.y, b.y, fr), sinus_lerp(a.z, b.z, fr), sinus_lerp(a.w, b.w, fr));
				if (ppbgOn) {
					const float4 g = __ldg(reinterpret_cast<const float4*>(ppbg) + j4);
					r.x = saturate01(r.x - fmaf(w, g.x, o)); r.y = saturate01(r.y - fmaf(w, g.y, o));
					r.z = saturate01(r.z - fmaf(w, g.z, o)); r.w = saturate01(r.w - fmaf(w, g.w, o));
				}
				*reinterpret_cast<float4*>(d + 4 * j4) = r;
				if constexpr (CONV)
					*reinterpret_cast<ushort4*>(c + 4 * j4) = make_ushort4(to_u16_container(r.x, convScale), to_u16_container(r.y, convScale),
					                                                       to_u16_container(r.z, convScale), to_u16_container(r.w, convScale));
			}
		} else {
			for (int j = lane; j < H; j += 32) {
				const float a = s0[j];
				float r = last ? a : sinus_lerp(a, s1[j], fr);
				if (ppbgOn) r = saturate01(r - fmaf(w, __ldg(ppbg + j), o));
				d[j] = r;
				if constexpr (CONV) c[j] = to_u16_container(r, convScale);
			}
		}
	}
}
cudaError_t launch_sinusoidal(float* out, const float* in, const float* curve, int H, int A, long long samples,
                              int ppbgOn, const float* ppbg, float w, float o, unsigned short* conv, float convScale,
                              int smCount, cudaStream_t st) {
	const long long lines64 = samples / H;
	if (lines64 < 1 || lines64 > 0x7fffffffLL) return cudaErrorInvalidValue;
	const int lines = (int)lines64;
	/* float4 / ushort4 path: line length a multiple of four and every base pointer 16-byte aligned (8 for the containers) */
	const uintptr_t al = reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in) | (ppbgOn ? reinterpret_cast<uintptr_t>(ppbg) : 0);
	const int vec = (H % 4 == 0) && (al & 15) == 0 && (reinterpret_cast<uintptr_t>(conv) & 7) == 0;
	long long blocks = ((long long)lines + 7) / 8;
	if (blocks > (long long)smCount * 8) blocks = (long long)smCount * 8;
	if (conv) sinusoidal_kernel<true><<<(int)blocks, 256, 0, st>>>(out, in, curve, H, A, lines, vec, ppbgOn, ppbg, w, o, conv, convScale);
	else sinusoidal_kernel<false><<<(int)blocks, 256, 0, st>>>(out, in, curve, H, A, lines, vec, ppbgOn, ppbg, w, o, nullptr, 0.f);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ post-process background */
__global__ void ppbg_record_kernel(float* __restrict__ bg, const float* __restrict__ data, int H, int A) {
	const int z = blockIdx.x * blockDim.x + threadIdx.x;
	if (z < H) {
		float sum = 0.f;
		for (int i = 0; i < A; ++i) sum += data[z + (size_t)i * H];
		bg[z] = __fdividef(sum, (float)A);
	}
}
cudaError_t launch_ppbg_record(float* bg, const float* data, int H, int A, cudaStream_t st) {
	ppbg_record_kernel<<<(H + 127) / 128, 128, 0, st>>>(bg, data, H, A);
	return cudaGetLastError();
}
__global__ void __launch_bounds__(256) ppbg_remove_kernel(float* __restrict__ data, const float* __restrict__ bg, float w, float o,
                                                          int H, long long samples) {
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < samples; i += (long long)gridDim.x * blockDim.x)
		data[i] = saturate01(data[i] - fmaf(w, __ldg(bg + (int)(i % H)), o));
}
cudaError_t launch_ppbg_remove(float* data, const float* bg, float w, float o, int H, long long samples, int smCount, cudaStream_t st) {
	long long blocks = (samples + 255) / 256;
	if (blocks > (long long)smCount * 16) blocks = (long long)smCount * 16;
	ppbg_remove_kernel<<<(int)blocks, 256, 0, st>>>(data, bg, w, o, H, samples);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ display extraction */
__global__ void bscan_frame_kernel(float* __restrict__ disp, const float* __restrict__ vol, unsigned Btot, unsigned F,
                                   unsigned frameNr, unsigned nFrames, int fn) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= F) return;
	if (nFrames > 1) {
		if (fn == 0) {
			int cnt = 0; float sum = 0.f;
			for (unsigned j = 0; j < nFrames; ++j) { const unsigned f = frameNr + j; if (f < Btot) { sum += vol[(size_t)f * F + (F - 1) - i]; ++cnt; } }
			disp[i] = __fdividef(sum, (float)cnt);
		} else if (fn == 1) {
			float mx = 0.f;
			for (unsigned j = 0; j < nFrames; ++j) { const unsigned f = frameNr + j; if (f < Btot) { const float v = vol[(size_t)f * F + (F - 1) - i]; if (mx < v) mx = v; } }
			disp[i] = mx;
		}
	} else {
		disp[i] = vol[(size_t)frameNr * F + (F - 1) - i];
	}
}
cudaError_t launch_bscan_frame(float* disp, const float* vol, unsigned Btot, unsigned F, unsigned frameNr, unsigned nFrames, int fn, cudaStream_t st) {
	bscan_frame_kernel<<<(F + 255) / 256, 256, 0, st>>>(disp, vol, Btot, F, frameNr, nFrames, fn);
	return cudaGetLastError();
}

/* en-face: one thread per lateral position i (A-scan of the volume); reads nFrames consecutive depth bins of its line */
__global__ void enface_frame_kernel(float* __restrict__ disp, const float* __restrict__ vol, unsigned W, unsigned E,
                                    unsigned frameNr, unsigned nFrames, int fn) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= E) return;
	if (nFrames > 1) {
		if (fn == 0) {
			int cnt = 0; float sum = 0.f;
			for (unsigned j = 0; j < nFrames; ++j) { const unsigned f = frameNr + j; if (f < W) { sum += vol[f + (size_t)i * W]; ++cnt; } }
			disp[(E - 1) - i] = __fdividef(sum, (float)cnt);
		} else if (fn == 1) {
			float mx = 0.f;
			for (unsigned j = 0; j < nFrames; ++j) { const unsigned f = frameNr + j; if (f < W) { const float v = vol[f + (size_t)i * W]; if (mx < v) mx = v; } }
			disp[(E - 1) - i] = mx;
		}
	} else {
		disp[(E - 1) - i] = vol[frameNr + (size_t)i * W];
	}
}
cudaError_t launch_enface_frame(float* disp, const float* vol, unsigned W, unsigned E, unsigned frameNr, unsigned nFrames, int fn, cudaStream_t st) {
	enface_frame_kernel<<<(E + 255) / 256, 256, 0, st>>>(disp, vol, W, E, frameNr, nFrames, fn);
	return cudaGetLastError();
}

/* ---- en-face extraction fused with its all-gather over peer memory (multi-GPU shards) ----
 * Same per-line arithmetic as enface_frame_kernel (cuda_code.cu:884-912); instead of a local frame + ncclAllGather every
 * value is stored straight into the frame window of EVERY rank (P2P stores over NVLink / NVSwitch; the own rank is a plain
 * store).  Protocol as in the fused kernel (oct_device.cuh GatherDev): acknowledgements of the frame buffer being overwritten are
 * awaited first, the last CTA to finish publishes `seq` in each rank's arrived[] word for this rank (release, system scope). */

__global__ void __launch_bounds__(256) enface_gather_kernel(const EnfaceGatherArgs a) {
	if (a.world > 0 && a.seq > (unsigned)OCT_GATHER_FRAMES) {
		if (threadIdx.x == 0) {
			const unsigned* acks = a.flags[a.rank] + OCT_GATHER_ACK;
			bool ok = true;
			for (int c = 0; c < a.world; ++c) ok = gather_spin_ge(acks + c, a.seq - (unsigned)OCT_GATHER_FRAMES) && ok;
			if (!ok) atomicAdd(a.status, 1u);
		}
		__syncthreads();
	}
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.E) {
		float val;
		if (a.nFrames > 1) {
			if (a.fn == 0) {
				int cnt = 0; float sum = 0.f;
				for (unsigned j = 0; j < a.nFrames; ++j) { const unsigned f = a.frameNr + j; if (f < a.W) { sum += a.vol[f + (size_t)i * a.W]; ++cnt; } }
				val = __fdividef(sum, (float)cnt);
			} else {
				float mx = 0.f;
				for (unsigned j = 0; j < a.nFrames; ++j) { const unsigned f = a.frameNr + j; if (f < a.W) { const float v = a.vol[f + (size_t)i * a.W]; if (mx < v) mx = v; } }
				val = mx;
			}
		} else {
			val = a.vol[a.frameNr + (size_t)i * a.W];
		}
		const unsigned dst = (a.Eglobal - 1u) - (a.offset + i);         /* the reference writes the frame reversed (cuda_code.cu:909) */
#pragma unroll 1
		for (int r = 0; r < a.world; ++r) a.frames[r][dst] = val;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		const unsigned done = atomicAdd(a.counter, 1u);
		if (done == gridDim.x - 1) {
			*a.counter = 0;                                             /* next launch is stream ordered behind this one */
			__threadfence_system();
#pragma unroll 1
			for (int r = 0; r < a.world; ++r) st_relaxed_sys_u32(a.flags[r] + OCT_GATHER_ARRIVED + a.rank, a.seq);
		}
	}
}
/* consumer side: wait until every rank's slab of frame `seq` has arrived, copy the frame out of the window into this rank's private
 * display frame (what the GL widget's PBO is in the reference), and only then acknowledge `seq` to every producer -- the producers
 * check that acknowledgement before they overwrite this frame buffer two gathers later (flow control: a rank that runs ahead can
 * never tear a frame a slower rank is still reading). */
__global__ void __launch_bounds__(256) enface_consume_kernel(const EnfaceConsumeArgs a) {
	/* in stream order behind the kernel that produced this rank's slab, launched as its programmatic dependent: the launch latency hides
	 * behind that kernel's tail, and the next compute kernel (this kernel's dependent) may run its prologue meanwhile */
	grid_launch_dependents();
	grid_dependency_wait();
	if ((int)threadIdx.x < a.world) {
		if (!gather_spin_ge(a.window + OCT_GATHER_ARRIVED + threadIdx.x, a.seq) && blockIdx.x == 0) atomicAdd(a.status + 1, 1u);
	}
	__syncthreads();
	const unsigned quads = a.Eglobal >> 2;
	const float4* src4 = reinterpret_cast<const float4*>(a.frame);
	float4* dst4 = reinterpret_cast<float4*>(a.display);
	for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) dst4[q] = __ldcg(src4 + q);
	for (unsigned i = (quads << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < a.Eglobal; i += gridDim.x * blockDim.x) a.display[i] = __ldcg(a.frame + i);
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const unsigned done = atomicAdd(a.counter, 1u);
		if (done == gridDim.x - 1) {
			*a.counter = 0;
			__threadfence_system();
#pragma unroll 1
			for (int r = 0; r < a.world; ++r) st_relaxed_sys_u32(a.peerHeaders[r] + OCT_GATHER_ACK + a.rank, a.seq);
		}
	}
}
cudaError_t launch_enface_gather(const EnfaceGatherArgs& a, cudaStream_t st) {
	enface_gather_kernel<<<(a.E + 255) / 256, 256, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t launch_enface_consume(const EnfaceConsumeArgs& a, int smCount, bool dependent, cudaStream_t st) {
	int blocks = (int)((a.Eglobal / 4 + 255) / 256);
	if (blocks > smCount) blocks = smCount;
	if (blocks < 1) blocks = 1;
	if (dependent) {
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at; cfg.numAttrs = 1;
		return cudaLaunchKernelEx(&cfg, enface_consume_kernel, a);
	}
	enface_consume_kernel<<<blocks, 256, 0, st>>>(a);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ dispersion sweep: A-scan sharpness metric per trial
 * AscanMetricCalculator::calculateMetric (octproz-dispersion-estimator-extension/src/ascanmetriccalculator.cpp:22-128) on the
 * [trials][lines][H] outputs of one sweep launch.  One block per trial, one thread per line; every sum runs in the reference's
 * order (samples ascending inside a line, then lines ascending), so the value is the reference's for the same processed data. */
__global__ void __launch_bounds__(128) sweep_metric_kernel(float* __restrict__ metrics, const float* __restrict__ data, int lines, int H,
                                                          int metric, float thr, int ignore) {
	extern __shared__ float lineMetric[];
	const float* trial = data + (size_t)blockIdx.x * lines * H;
	int ig = ignore;
	if (ig > 0) ig = ig < H ? ig : H;                      /* clamp, ascanmetriccalculator.cpp:39-41 */
	const int valid = H - ig;
	for (int l = threadIdx.x; l < lines; l += blockDim.x) {
		const float* d = trial + (size_t)l * H + ig;
		float m = 0.0f;
		if (valid > 0) {
			if (metric == 0) { for (int i = 0; i < valid; ++i) { const float v = d[i]; if (v > thr) m += v; } }
			else if (metric == 1) { int c = 0; for (int i = 0; i < valid; ++i) if (d[i] > thr) ++c; m = (float)c; }
			else if (metric == 2) { for (int i = 0; i < valid; ++i) m = fmaxf(m, d[i]); }
			else if (metric == 3 && valid >= 3) {
				float s = 0.0f; int c = 0;
				for (int i = 1; i < valid - 1; ++i) { s += fabsf((d[i + 1] - d[i - 1]) * 0.5f); ++c; }
				m = c > 0 ? s / (float)c : 0.0f;
			}
		}
		lineMetric[l] = (valid > 0) ? m : 0.0f;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		float total = 0.0f;
		for (int l = 0; l < lines; ++l) total += lineMetric[l];
		metrics[blockIdx.x] = total;
	}
}
cudaError_t launch_sweep_metric(float* metrics, const float* data, int trials, int lines, int H, int metric, float thr, int ignore, cudaStream_t st) {
	sweep_metric_kernel<<<trials, 128, (size_t)lines * sizeof(float), st>>>(metrics, data, lines, H, metric, thr, ignore);
	return cudaGetLastError();
}
/* ------------------------------------------------------------------ 12-bit packed -> u16 containers
 * little-endian bit packing (GenICam Mono12p): sample k of an octet lives in bits [12k, 12k+12) of its 96 bits */
__global__ void __launch_bounds__(256) unpack12_kernel(uint4* __restrict__ out, const unsigned* __restrict__ in, long long octets) {
	for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < octets; o += (long long)gridDim.x * blockDim.x) {
		const unsigned a0 = __ldg(in + 3 * o), a1 = __ldg(in + 3 * o + 1), a2 = __ldg(in + 3 * o + 2);
		const unsigned s0 = a0 & 0xFFFu, s1 = (a0 >> 12) & 0xFFFu, s2 = __funnelshift_r(a0, a1, 24) & 0xFFFu, s3 = (a1 >> 4) & 0xFFFu;
		const unsigned s4 = (a1 >> 16) & 0xFFFu, s5 = __funnelshift_r(a1, a2, 28) & 0xFFFu, s6 = (a2 >> 8) & 0xFFFu, s7 = a2 >> 20;
		out[o] = make_uint4(s0 | (s1 << 16), s2 | (s3 << 16), s4 | (s5 << 16), s6 | (s7 << 16));
	}
}
cudaError_t launch_unpack12(uint16_t* out, const void* in, long long octets, int smCount, cudaStream_t st) {
	long long blocks = (octets + 255) / 256;
	if (blocks > (long long)smCount * 16) blocks = (long long)smCount * 16;
	unpack12_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(out), static_cast<const unsigned*>(in), octets);
	return cudaGetLastError();
}

/* u8 voxels in the layout of the GL_R8 3-D texture (x = A-scan, y = B-scan in volume, z flipped depth), cuda_code.cu:928-940.
 * The slab is [B-scan][A-scan][depth] with depth fastest, the texture [depth][B-scan][A-scan] with the A-scan fastest: a transpose of
 * every B-scan.  The reference (and the first version here) wrote one byte per thread with a stride of A * Btot bytes -- every byte
 * its own 32-byte sector.  Here a block moves a 64 (A-scans) x 64 (depth) tile through shared memory: coalesced 128-byte reads along
 * the depth, conversion, 64-byte row writes along the A-scan axis. */
constexpr int VT = 64, VPITCH = 68;
__global__ void __launch_bounds__(256) volume_u8_kernel(uint8_t* __restrict__ tex, const float* __restrict__ buf,
                                                         unsigned bufferNr, unsigned B, unsigned A, unsigned Btot, unsigned depth,
                                                         unsigned tilesA, unsigned tilesZ, unsigned tiles, int vec) {
	__shared__ __align__(16) unsigned char tile[VT * VPITCH];
	const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (unsigned t = blockIdx.x; t < tiles; t += gridDim.x) {
		const unsigned tz = t % tilesZ, ta = (t / tilesZ) % tilesA, bl = t / (tilesZ * tilesA);
		const unsigned a0 = ta * VT, z0 = tz * VT;
		/* load + convert: warp w takes A-scans a0 + w, a0 + w + 8, ...; lanes run along the depth */
#pragma unroll
		for (unsigned r = 0; r < VT / 8; ++r) {
			const unsigned al = warp + 8 * r, a = a0 + al;
			if (a < A) {
				const float* line = buf + ((size_t)bl * A + a) * depth;
#pragma unroll
				for (unsigned h = 0; h < 2; ++h) {
					const unsigned zl = lane + 32 * h, z = z0 + zl;
					if (z < depth) tile[zl * VPITCH + al] = (unsigned char)((double)line[z] * 255.0);     /* cuda_code.cu:935 */
				}
			}
		}
		__syncthreads();
		/* store: texture row (z', y) holds A bytes; 16 threads write one 64-byte row segment */
		const unsigned y = bl + bufferNr * B;
#pragma unroll
		for (unsigned it = 0; it < 4; ++it) {
			const unsigned zl = (threadIdx.x >> 4) + 16 * it, z = z0 + zl, c = (threadIdx.x & 15) * 4, a = a0 + c;
			if (z < depth && a < A) {
				unsigned char* dst = tex + ((size_t)(depth - 1 - z) * Btot + y) * A + a;
				if (vec && a + 3 < A) *reinterpret_cast<unsigned*>(dst) = *reinterpret_cast<const unsigned*>(tile + zl * VPITCH + c);
				else for (unsigned k = 0; k < 4 && a + k < A; ++k) dst[k] = tile[zl * VPITCH + c + k];
			}
		}
		__syncthreads();
	}
}
cudaError_t launch_volume_u8(uint8_t* tex, const float* buf, long long samples, unsigned bufferNr, unsigned B, unsigned A, unsigned Btot,
                             unsigned depth, int smCount, cudaStream_t st) {
	if (samples != (long long)B * A * depth) return cudaErrorInvalidValue;
	const unsigned tilesA = (A + VT - 1) / VT, tilesZ = (depth + VT - 1) / VT;
	const unsigned long long tiles64 = (unsigned long long)tilesA * tilesZ * B;
	if (tiles64 == 0 || tiles64 > 0x7fffffffULL) return cudaErrorInvalidValue;
	const unsigned tiles = (unsigned)tiles64;
	const int vec = (A % 4 == 0) && (reinterpret_cast<uintptr_t>(tex) & 3) == 0;
	unsigned blocks = tiles < (unsigned)smCount * 8u ? tiles : (unsigned)smCount * 8u;
	volume_u8_kernel<<<blocks, 256, 0, st>>>(tex, buf, bufferNr, B, A, Btot, depth, tilesA, tilesZ, tiles, vec);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------ float -> output container */
__global__ void __launch_bounds__(256) float_to_output_kernel(void* __restrict__ out, const float* __restrict__ in, int bitDepth, long long n) {
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
		const float s = __saturatef(in[i]);
		if (bitDepth <= 8) reinterpret_cast<unsigned char*>(out)[i] = (unsigned char)((double)s * 255.0);
		else if (bitDepth <= 10) reinterpret_cast<unsigned short*>(out)[i] = (unsigned short)((double)s * 1023.0);
		else if (bitDepth <= 12) reinterpret_cast<unsigned short*>(out)[i] = (unsigned short)((double)s * 4095.0);
		else if (bitDepth <= 16) reinterpret_cast<unsigned short*>(out)[i] = (unsigned short)((double)s * 65535.0);
		else if (bitDepth <= 24) reinterpret_cast<unsigned int*>(out)[i] = (unsigned int)(s * 16777215.0f);
		else reinterpret_cast<unsigned int*>(out)[i] = (unsigned int)(s * 4294967295.0f);
	}
}
cudaError_t launch_float_to_output(void* out, const float* in, int bitDepth, long long n, int smCount, cudaStream_t st) {
	long long blocks = (n + 255) / 256;
	if (blocks > (long long)smCount * 16) blocks = (long long)smCount * 16;
	float_to_output_kernel<<<(int)blocks, 256, 0, st>>>(out, in, bitDepth, n);
	return cudaGetLastError();
}

}  // namespace octb200
