/*
 * emu_phases.cpp -- TEST-ONLY lane emulator.  Executes the __host__ __device__ phases of
 * octproz_b200/csrc/oct_phases.cuh for all 32 lanes (and both warps of an R=2 line group) on the CPU so
 * the register / lane / shared-memory index maps of the fused kernel can be checked without a GPU.
 * Never linked into the product library.
 */
#include "oct_luts.hpp"
#include <vector>

using namespace octb200;

extern "C" int emu_fft32(const float* in, float* out) {
	float2 v[32];
	for (int i = 0; i < 32; ++i) v[i] = make_float2(in[2 * i], in[2 * i + 1]);
	fft32_inv(v);
	for (int r = 0; r < 32; ++r) { out[2 * bitrev5(r)] = v[r].x; out[2 * bitrev5(r) + 1] = v[r].y; }
	return 0;
}

/* one 1024-point sub transform for all lanes: in[s] (s<1024) -> regs[lane][r] = E[lane + 32*bitrev5(r)] */
static void sub_fft_1024(const std::vector<float2>& x, std::vector<float2>& tile, float2 (*regs)[32],
                         const std::vector<float2>& tw) {
	for (int lane = 0; lane < 32; ++lane) {
		float2 (&v)[32] = regs[lane];
		for (int j = 0; j < 32; ++j) v[j] = x[lane + 32 * j];
		fft32_inv(v);
		exchange_store(lane, v, tile.data(), tw.data());
	}
	for (int lane = 0; lane < 32; ++lane) {
		float2 (&v)[32] = regs[lane];
		exchange_load(lane, v, tile.data());
		fft32_inv(v);
	}
}

extern "C" int emu_ifft_1024(const float* in, float* out /* 1024 complex, natural order */) {
	std::vector<float2> x(1024), tile(XBUF_FLOAT2), tw;
	build_twiddles_1024(tw);
	for (int i = 0; i < 1024; ++i) x[i] = make_float2(in[2 * i], in[2 * i + 1]);
	static float2 regs[32][32];
	sub_fft_1024(x, tile, regs, tw);
	for (int lane = 0; lane < 32; ++lane)
		for (int r = 0; r < 32; ++r) {
			const int k = lane + 32 * bitrev5(r);
			out[2 * k] = regs[lane][r].x; out[2 * k + 1] = regs[lane][r].y;
		}
	return 0;
}

/* full fused line: float slot -> stage A -> FFT -> epilogue.  R = 1 (N=1024) or 2 (N=2048). */
extern "C" int emu_fused_line(int N, int sa, int interp, const float* fslot /* halo-less line, N floats (+16 slack) */,
                              int shift, const float* resample, const float* window, const float* phasor /* N pairs or NULL */,
                              int logMode, float gmin, float gmax, float coeff, float addend,
                              const float* meanLine /* N/2 pairs or NULL */, const float* ppbg, float ppbgW, float ppbgO,
                              float* outLine /* N/2 */, float* cplxLine /* N/2 pairs or NULL */) {
	const int R = N / 1024;
	if (R != 1 && R != 2) return -1;
	std::vector<float4> lut;
	const bool four = (sa == SA_CUBIC || sa == SA_LINEAR);
	const bool split = stage_a_splits_slot(R, sa, false);
	build_stage_luts_paired(N, R, interp, resample, window, reinterpret_cast<const float2*>(phasor), lut, four ? (split ? 2 : 1) : 0);
	std::vector<float2> tw, ctw;
	build_twiddles_1024(tw);
	build_combine_twiddles_2048(ctw);
	EpiConsts e = make_epi_consts(N, logMode, gmin, gmax, coeff, addend);
	e.fpn = meanLine != nullptr; e.ppbg = ppbg != nullptr; e.ppbgWeight = ppbgW; e.ppbgOffset = ppbgO;
	/* cubic: the kernel mirrors f[-1] = f[1]; the caller's buffer has room in front of the line */
	if (sa == SA_CUBIC) const_cast<float*>(fslot)[-1] = fslot[1];
	/* R = 2, 4-tap: the kernel's conversion writes the slot split by sample parity */
	std::vector<float> splitSlot;
	if (split) {
		splitSlot.assign(FSLOT_PAD + SPLIT_ODD_BASE + 1024 + 8, 0.0f);
		float* e = splitSlot.data() + FSLOT_PAD;
		for (int k = 0; k < 1024; ++k) { e[k] = fslot[2 * k]; e[SPLIT_ODD_BASE + k] = fslot[2 * k + 1]; }
		e[SPLIT_ODD_BASE - 1] = fslot[1];
		fslot = e;
	}
	static float2 regs[2][32][32];
	std::vector<float2> tile[2] = { std::vector<float2>(XBUF_FLOAT2), std::vector<float2>(XBUF_FLOAT2) };
	for (int p = 0; p < R; ++p) {
		for (int lane = 0; lane < 32; ++lane) {
			float2 (&v)[32] = regs[p][lane];
			const float4* B = lut.data();
			if (sa == SA_CUBIC) { if (R == 1) stage_a<SA_CUBIC, 1>(lane, p, fslot, shift, B, v); else stage_a<SA_CUBIC, 2>(lane, p, fslot, shift, B, v); }
			else if (sa == SA_LINEAR) { if (R == 1) stage_a<SA_LINEAR, 1>(lane, p, fslot, shift, B, v); else stage_a<SA_LINEAR, 2>(lane, p, fslot, shift, B, v); }
			else if (sa == SA_NONE) { if (R == 1) stage_a<SA_NONE, 1>(lane, p, fslot, shift, B, v); else stage_a<SA_NONE, 2>(lane, p, fslot, shift, B, v); }
			else { if (R == 1) stage_a<SA_LANCZOS, 1>(lane, p, fslot, shift, B, v); else stage_a<SA_LANCZOS, 2>(lane, p, fslot, shift, B, v); }
			if (stage_a_fuses_stage0(sa)) fft32_inv<true>(v); else fft32_inv(v);
			exchange_store(lane, v, tile[p].data(), tw.data());
		}
		for (int lane = 0; lane < 32; ++lane) {
			float2 (&v)[32] = regs[p][lane];
			exchange_load(lane, v, tile[p].data());
			fft32_inv(v);
		}
	}
	const float2* mean = reinterpret_cast<const float2*>(meanLine);
	if (R == 1) {
		for (int lane = 0; lane < 32; ++lane) {
			if (cplxLine) epilogue_complex<0>(lane, regs[0][lane], reinterpret_cast<float2*>(cplxLine));
			epilogue_scaled<0>(lane, regs[0][lane], e, mean, ppbg, outLine);
		}
	} else {
		for (int p = 0; p < 2; ++p)
			for (int lane = 0; lane < 32; ++lane) combine_store(lane, p, regs[p][lane], tile[p].data(), ctw.data());
		for (int p = 0; p < 2; ++p)
			for (int lane = 0; lane < 32; ++lane) combine_load(lane, p, regs[p][lane], tile[1 - p].data());
		for (int lane = 0; lane < 32; ++lane) {
			if (cplxLine) {
				epilogue_complex<0>(lane, regs[0][lane], reinterpret_cast<float2*>(cplxLine));
				epilogue_complex<16>(lane, regs[1][lane], reinterpret_cast<float2*>(cplxLine));
			}
			epilogue_scaled<0>(lane, regs[0][lane], e, mean, ppbg, outLine);
			epilogue_scaled<16>(lane, regs[1][lane], e, mean, ppbg, outLine);
		}
	}
	return 0;
}

/* ---- the shared-memory kernel's transform (generic_fft.cuh): plan, per-pass twiddle tables, padded line buffers and the Stockham
 * passes exactly as oct_generic_kernel runs them, with T "threads" executed one after the other (a pass reads one buffer and writes the
 * other, so the order of the threads inside a pass does not matter -- as on the GPU between two barriers) ---- */
#include "generic_fft.cuh"

extern "C" int emu_generic_ifft(int N, int T, const float* in, float* out /* N complex, natural order */, int* radixOut, int* nPassOut) {
	int radix[16] = {}, nPass = 0, twOff[16]; unsigned magic[16];
	if (!generic_fft_plan(N, radix, &nPass)) return -1;
	const int entries = generic_twiddle_layout(N, radix, nPass, twOff, magic);
	std::vector<float2> tw((size_t)entries, make_float2(1.f, 0.f));
	generic_fill_twiddles(radix, nPass, twOff, tw.data());
	std::vector<float2> A((size_t)gpad_len(N)), B((size_t)gpad_len(N));
	for (int m = 0; m < N; ++m) A[gpad(m)] = make_float2(in[2 * m], in[2 * m + 1]);
	const float2* src = A.data(); float2* dst = B.data();
	int Ns = 1;
	for (int ps = 0; ps < nPass; ++ps) {
		const int R = radix[ps];
		for (int tid = 0; tid < T; ++tid) {
			switch (R) {
			case 2: stockham_pass<2>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 3: stockham_pass<3>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 4: stockham_pass<4>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 5: stockham_pass<5>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 7: stockham_pass<7>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 8: stockham_pass<8>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 11: stockham_pass<11>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			case 13: stockham_pass<13>(src, dst, N, Ns, magic[ps], tw.data() + twOff[ps], tid, T); break;
			default: return -2;
			}
		}
		Ns *= R;
		const float2* t = src; src = dst; dst = const_cast<float2*>(t);
	}
	for (int z = 0; z < N; ++z) { out[2 * z] = src[gpad(z)].x; out[2 * z + 1] = src[gpad(z)].y; }
	if (radixOut) for (int i = 0; i < nPass; ++i) radixOut[i] = radix[i];
	if (nPassOut) *nPassOut = nPass;
	return 0;
}
