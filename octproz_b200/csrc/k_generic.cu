/*
 * k_generic.cu -- the fused raw -> B-scan kernel for every line length and container the register kernel (k_fused.cuh) does not take:
 * any even N <= 8192 whose prime factors are in {2, 3, 5, 7, 11, 13} (the reference's default acquisition geometry is N = 1664 =
 * 2^7 * 13, octproz/default/settings.ini:62) and u8 / u16 / u32 containers.  Same stages, same per-sample arithmetic:
 *
 *   raw line --cp.async.bulk (TMA 1-D) + mbarrier--> shared memory                                   (one CTA works on one line at a time)
 *     -> container -> fp32 [+ rolling-mean background removal]                                        (cuda_code.cu:109-211)
 *     -> resampling x window x dispersion phasor from the natural-order stage LUT                     (cuda_code.cu:213-489)
 *     -> inverse FFT: Stockham autosort passes between two (padded) shared-memory line buffers, mixed radix
 *        (radix 8 / 4 / 2 butterflies, odd primes by the symmetric O(P^2/2) form), per-pass twiddle tables
 *        laid out so that a warp reads consecutive words                                                (cuda_code.cu:1514-1515)
 *     -> FPN subtract, |.|^2, log / linear scale, truncate to N/2, flip in the store address, background  (cuda_code.cu:567-807)
 *     -> coalesced fp32 stores.  HBM traffic: the container bytes in + 2 B out per raw sample; nothing in between.
 *
 * The transform lives in shared memory instead of registers: ~3x the shared-memory traffic and more instructions per sample than the
 * N = 1024 / 2048 register kernels, but one launch and 4 B/sample of HBM traffic instead of the 32 B/sample of the
 * pre-kernel + cuFFT + post-kernel chain, which remains the path for line lengths with larger prime factors.
 */
#include "k_aux.cuh"
#include "generic_fft.cuh"

#include <cstdlib>

namespace octb200 {

template <typename RawT>
__device__ __forceinline__ float generic_convert(RawT v, int shiftBits) {
	if constexpr (sizeof(RawT) == 4) {
		if (shiftBits) return (float)((double)v / 4294967296.0);   /* cuda_code.cu:144 */
		return __uint2float_rd(v);                                  /* cuda_code.cu:124 */
	} else {
		return __uint2float_rd((unsigned)v >> shiftBits);           /* cuda_code.cu:118-121,138-141 */
	}
}

/* shared memory of one CTA: [twiddle tables][per line team: buffer A, buffer B (aliased by the fp32 slot + rolling prefix sums)]
 * [raw region of the LB consecutive lines of a batch, with the Lanczos halos at both ends][mbarrier] */
struct GenericSmem { int twBytes, aBytes, bBytes, teamBytes, rawOff, rawBytes, barOff, total; };
__host__ __device__ inline GenericSmem generic_smem_layout(int N, int HB, int HA, int rawBytesPerSample, bool roll, int LB, int twEntries) {
	GenericSmem L;
	const int SE = HB + N + HA;
	L.twBytes = align_up(twEntries * 8, 128);
	L.aBytes = align_up(gpad_len(N) * 8, 128);
	int bBytes = gpad_len(N) * 8;
	const int slotBytes = align_up((FSLOT_PAD + SE) * 4, 16) + (roll ? align_up((SE + 1) * 8, 16) : 0);
	if (slotBytes > bBytes) bBytes = slotBytes;
	L.bBytes = align_up(bBytes, 128);
	L.teamBytes = L.aBytes + L.bBytes;
	L.rawOff = L.twBytes + LB * L.teamBytes;
	L.rawBytes = align_up((LB * N + HB + HA) * rawBytesPerSample, 128);
	L.barOff = L.rawOff + L.rawBytes;
	L.total = L.barOff + 128;
	return L;
}

/* blockDim = (TT, LB): LB "line teams" of TT threads; a CTA works on batches of LB consecutive lines (one bulk copy per batch), team l
 * on line l of the batch.  The teams share the twiddle tables (shared memory, filled once per CTA), the stage LUT reads (L1) and the
 * CTA-wide barriers between the passes. */
template <typename RawT, int SA, bool ROLL>
__global__ void __launch_bounds__(512, 2) oct_generic_kernel(const GenericArgs a) {
	extern __shared__ __align__(128) unsigned char smem[];
	constexpr int RB = sizeof(RawT);
	const int N = a.N, H = N / 2, SE = a.HB + N + a.HA;
	const int tid = threadIdx.x, T = blockDim.x, team = threadIdx.y, LB = blockDim.y;
	const int ctid = team * T + tid, CT = T * LB;
	const GenericSmem L = generic_smem_layout(N, a.HB, a.HA, RB, ROLL, LB, a.twEntries);
	float2* stw = reinterpret_cast<float2*>(smem);
	float2* bufA = reinterpret_cast<float2*>(smem + L.twBytes + team * L.teamBytes);
	float2* bufB = reinterpret_cast<float2*>(smem + L.twBytes + team * L.teamBytes + L.aBytes);
	float* fslot = reinterpret_cast<float*>(bufB) + FSLOT_PAD;
	unsigned long long* prefix = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(bufB) + align_up((FSLOT_PAD + SE) * 4, 16));
	RawT* rawRegion = reinterpret_cast<RawT*>(smem + L.rawOff);
	const RawT* rslot = rawRegion + team * N;                  /* this team's window [line start - HB, line end + HA) of the region */
	uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.barOff);
	const RawT* raw = reinterpret_cast<const RawT*>(a.raw);

	auto issue = [&](int g) {                /* batch of lines g .. g + LB - 1: one thread (bulk) or all threads (plain loads) */
		const int nl = min(LB, a.lines - g);
		const long long lo = (long long)g * N - a.HB, hi = (long long)(g + nl) * N + a.HA;
		const long long clo = lo < 0 ? 0 : lo, chi = hi > a.totalSamples ? a.totalSamples : hi;
		if (a.useBulk) {
			if (ctid == 0) {
				for (long long q = 0; q < clo - lo; ++q) rawRegion[q] = 0;
				for (long long q = chi - lo; q < hi - lo; ++q) rawRegion[q] = 0;
				const uint32_t bytes = (uint32_t)((chi - clo) * RB);
				mbar_arrive_expect_tx(bar, bytes);
				bulk_g2s(rawRegion + (clo - lo), raw + clo, bytes, bar);
			}
		} else {
			for (long long q = lo + ctid; q < hi; q += CT) rawRegion[q - lo] = (q >= 0 && q < a.totalSamples) ? raw[q] : (RawT)0;
		}
	};

	if (ctid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
	for (int i = ctid; i < a.twEntries; i += CT) stw[i] = __ldg(a.tw + i);
	__syncthreads();
	const int g0 = blockIdx.x * LB, gstep = gridDim.x * LB;
	if (g0 < a.lines) issue(g0);

	int it = 0;
	for (int g = g0; g < a.lines; g += gstep, ++it) {
		const int gline = g + team;
		const bool active = gline < a.lines;
		if (a.useBulk) mbar_wait(bar, (uint32_t)(it & 1));
		else __syncthreads();

		/* ---- container -> fp32 (cuda_code.cu:109-147), rolling-mean background (cuda_code.cu:165-211: exact integer prefix sums) ---- */
		if (active) {
			if constexpr (ROLL) {
				if (tid < 32) {
					unsigned long long carry = 0;
					if (tid == 0) prefix[0] = 0;
					for (int c = 0; c < SE; c += 32) {
						const int q = c + tid;
						unsigned long long x = 0;
						if (q < SE) x = (sizeof(RawT) == 4) ? (unsigned long long)rslot[q] : (unsigned long long)((unsigned)rslot[q] >> a.shiftBits);
#pragma unroll
						for (int d = 1; d < 32; d <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d); if (tid >= d) x += y; }
						if (q < SE) prefix[q + 1] = carry + x;
						carry += __shfl_sync(0xffffffffu, x, 31);
					}
				}
			}
			for (int q = tid; q < SE; q += T) fslot[q] = generic_convert<RawT>(rslot[q], a.shiftBits);
		}
		__syncthreads();
		/* raw region consumed: start the load of this CTA's next batch */
		if (g + gstep < a.lines) issue(g + gstep);
		if constexpr (ROLL) {
			if (active) {
				const int W = a.W;
				for (int q = tid; q < SE; q += T) {
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
			}
			__syncthreads();
		}
		if constexpr (SA == SA_CUBIC) {
			if (active && tid == 0) fslot[a.HB - 1] = fslot[a.HB + 1];      /* mirrored first tap of the cubic (cuda_code.cu:284) */
			__syncthreads();
		}

		/* ---- stage A: resampling x window x phasor -> complex FFT input in buffer A ---- */
		if (active) {
			const float* f = fslot + a.HB;
			const int shift = (SA == SA_LANCZOS && gline == 0) ? 8 : 0;
#pragma unroll 2
			for (int m = tid; m < N; m += T) {
				float2 val;
				if constexpr (SA == SA_CUBIC) {
					/* 4-tap interpolators in tap-weight form (Catmull-Rom cuda_code.cu:258-271 expanded per tap, or linear :229 as (0, 1-t, t, 0)):
					 * the same arithmetic as the register kernels, y = sum_k w_k f[n1 - 1 + k], then y * (window * phasor) */
					const float4 G0 = __ldg(a.lutG + 2 * m), G1 = __ldg(a.lutG + 2 * m + 1);
					const int o = __float_as_int(G0.x);
					const float y = fmaf(G1.x, ldf(f, o + 12), fmaf(G0.w, ldf(f, o + 8), fmaf(G0.z, ldf(f, o + 4), G0.y * ldf(f, o))));
					val = cscale(make_float2(G1.y, G1.z), y);
				} else {
					const float4 B = __ldg(a.lutB + m);
					if constexpr (SA == SA_NONE) val = sample_none(f, m, B);
					else val = sample_lanczos(f, shift, B);
				}
				bufA[gpad(m)] = val;
			}
		}
		__syncthreads();

		/* ---- inverse FFT: Stockham passes A -> B -> A ... ---- */
		const float2* in = bufA; float2* out = bufB;
		int Ns = 1;
		for (int ps = 0; ps < a.nPass; ++ps) {
			const int R = a.radix[ps];
			const unsigned magic = a.magic[ps];
			const float2* tw = stw + a.twOff[ps];
			if (active) {
				switch (R) {
				case 2: stockham_pass<2>(in, out, N, Ns, magic, tw, tid, T); break;
				case 3: stockham_pass<3>(in, out, N, Ns, magic, tw, tid, T); break;
				case 4: stockham_pass<4>(in, out, N, Ns, magic, tw, tid, T); break;
				case 5: stockham_pass<5>(in, out, N, Ns, magic, tw, tid, T); break;
				case 7: stockham_pass<7>(in, out, N, Ns, magic, tw, tid, T); break;
				case 8: stockham_pass<8>(in, out, N, Ns, magic, tw, tid, T); break;
				case 11: stockham_pass<11>(in, out, N, Ns, magic, tw, tid, T); break;
				default: stockham_pass<13>(in, out, N, Ns, magic, tw, tid, T); break;
				}
			}
			Ns *= R;
			__syncthreads();
			const float2* t = in; in = out; out = const_cast<float2*>(t);
		}

		/* ---- epilogue (bins z < N/2), from `in` = the buffer the last pass wrote ---- */
		if (active) {
			if (a.cplxOut != nullptr) {
				float2* o = a.cplxOut + (size_t)gline * H;
				for (int z = tid; z < H; z += T) o[z] = in[gpad(z)];
			} else {
				int b = gline / a.A, al = gline - b * a.A;
				if (a.flip && (((unsigned)b + a.bscanBase) & 1u) == 0u && (unsigned)b + a.bscanBase < a.flipEnd) al = a.A - 1 - al;
				float* o = a.out + ((size_t)b * a.A + al) * H;
				const EpiConsts e = a.epi;
				for (int z = tid; z < H; z += T) {
					float2 d = in[gpad(z)];
					if (e.fpn) d = csub(d, __ldg(a.meanLine + z));
					const float pw = fmaf(d.x, d.x, d.y * d.y);
					float v = e.logMode ? fmaf(oct_lg2(pw), e.scaleA, e.scaleB) : fmaf(oct_sqrt(pw), e.scaleA, e.scaleB);
					if (e.ppbg) v = saturate01(v - fmaf(e.ppbgWeight, __ldg(a.ppbg + z), e.ppbgOffset));
					o[z] = v;
				}
			}
		}
		__syncthreads();          /* both line buffers are free again before the next batch's conversion writes into B */
	}
}

/* CTA shape: TT threads per line by the line length; as many line teams as fit a 512-thread CTA and ~110 KB of shared memory (two CTAs
 * per SM), at least one.  OCTB200_GENERIC_LB overrides the team count (experiments). */
static void generic_shape(int N, int HB, int HA, int rawBytes, bool roll, int twEntries, int* TT, int* LB, int* smemBytes) {
	int tt = N >= 2048 ? 256 : (N >= 1024 ? 128 : 64);
	const char* envT = getenv("OCTB200_GENERIC_TT");       /* experiments */
	if (envT && (atoi(envT) == 64 || atoi(envT) == 128 || atoi(envT) == 256 || atoi(envT) == 512)) tt = atoi(envT);
	int lb = 1;        /* one line per CTA: more teams per CTA bought nothing on a B200 (profiles/r02f_generic_lb_sweep.txt) */
	const char* env = getenv("OCTB200_GENERIC_LB");
	if (env && atoi(env) > 0 && atoi(env) * tt <= 512) lb = atoi(env);
	else while (lb > 1 && generic_smem_layout(N, HB, HA, rawBytes, roll, lb, twEntries).total > 112 * 1024) lb >>= 1;
	while (lb > 1 && generic_smem_layout(N, HB, HA, rawBytes, roll, lb, twEntries).total > 227 * 1024) lb >>= 1;
	*TT = tt; *LB = lb; *smemBytes = generic_smem_layout(N, HB, HA, rawBytes, roll, lb, twEntries).total;
}

template <typename RawT, int SA, bool ROLL>
static cudaError_t launch_generic_t(const GenericArgs& a, int smCount, cudaStream_t st) {
	int TT = 0, LB = 0, smem = 0;
	generic_shape(a.N, a.HB, a.HA, (int)sizeof(RawT), ROLL, a.twEntries, &TT, &LB, &smem);
	if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
	auto k = oct_generic_kernel<RawT, SA, ROLL>;
	cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if (e != cudaSuccess) return e;
	int ctasPerSm = (227 * 1024) / (smem + 1024);
	const int maxByThreads = 2048 / (TT * LB);
	if (ctasPerSm > maxByThreads) ctasPerSm = maxByThreads;
	if (ctasPerSm < 1) ctasPerSm = 1;
	int grid = smCount * ctasPerSm;
	const int batches = (a.lines + LB - 1) / LB;
	if (grid > batches) grid = batches;
	if (grid < 1) grid = 1;
	k<<<grid, dim3(TT, LB), smem, st>>>(a);
	return cudaGetLastError();
}

template <typename RawT>
static cudaError_t launch_generic_raw(const GenericArgs& a, int sa, bool roll, int smCount, cudaStream_t st) {
	/* linear = the same 4-tap kernel with weights (0, 1-t, t, 0) */
	if (sa == SA_CUBIC || sa == SA_LINEAR) return roll ? launch_generic_t<RawT, SA_CUBIC, true>(a, smCount, st) : launch_generic_t<RawT, SA_CUBIC, false>(a, smCount, st);
	if (sa == SA_NONE) return roll ? launch_generic_t<RawT, SA_NONE, true>(a, smCount, st) : launch_generic_t<RawT, SA_NONE, false>(a, smCount, st);
	return roll ? launch_generic_t<RawT, SA_LANCZOS, true>(a, smCount, st) : launch_generic_t<RawT, SA_LANCZOS, false>(a, smCount, st);
}

cudaError_t launch_generic(const GenericArgs& a, int rawBytes, int sa, bool roll, int smCount, cudaStream_t st) {
	if (rawBytes == 1) return launch_generic_raw<uint8_t>(a, sa, roll, smCount, st);
	if (rawBytes == 2) return launch_generic_raw<uint16_t>(a, sa, roll, smCount, st);
	return launch_generic_raw<uint32_t>(a, sa, roll, smCount, st);
}

bool generic_fits(int N, int rawBytes, int HB, int HA, bool roll, int twEntries) {
	return generic_smem_layout(N, HB, HA, rawBytes, roll, 1, twEntries).total <= 227 * 1024;
}

}  // namespace octb200
