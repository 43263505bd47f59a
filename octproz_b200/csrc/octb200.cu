/*
 * octb200.cu -- the C ABI (include/octb200.h): host orchestration of the B200 pipeline.
 *
 * Replaces, behind a handle, the file-scope-global state machine of the reference
 * (cuda_code.cu:39-105 globals, :1067-1162 initializeCuda, :1389-1605 octCudaPipeline,
 *  :1164-1212 cleanup).  Differences by design: per-handle state (re-entrant), status codes
 * instead of exit(), per-slot raw buffers and events instead of eight round-robin streams sharing
 * one set of buffers (latent cross-stream hazard), one fused kernel instead of 6-12.
 */
#include "../../include/octb200.h"
#include "k_aux.cuh"
#include "generic_fft.cuh"
#include "oct_curves.hpp"
#include "oct_luts.hpp"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace octb200;

namespace {

thread_local std::string g_createError;   /* octb200_last_error(NULL): the failure of THIS thread's last octb200_create */

/* ---- lazily bound cuFFT (only OCTB200_FFT_CUFFT touches it; the measured library baseline) ---- */
struct CufftApi {
	void* lib = nullptr;
	int (*Plan1d)(int*, int, int, int) = nullptr;
	int (*SetStream)(int, cudaStream_t) = nullptr;
	int (*ExecC2C)(int, float2*, float2*, int) = nullptr;
	int (*Destroy)(int) = nullptr;
	bool load(std::string& err) {
		if (lib) return true;
		const char* names[] = { "libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so" };
		for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
		if (!lib) { err = "cannot load libcufft.so.11"; return false; }
		Plan1d = (int (*)(int*, int, int, int))dlsym(lib, "cufftPlan1d");
		SetStream = (int (*)(int, cudaStream_t))dlsym(lib, "cufftSetStream");
		ExecC2C = (int (*)(int, float2*, float2*, int))dlsym(lib, "cufftExecC2C");
		Destroy = (int (*)(int))dlsym(lib, "cufftDestroy");
		if (!Plan1d || !SetStream || !ExecC2C || !Destroy) { err = "libcufft symbols missing"; return false; }
		return true;
	}
};
CufftApi g_cufft;
constexpr int kCufftC2C = 0x29, kCufftInverse = 1;

int rup(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

struct octb200_pipeline {
	octb200_config cfg{};
	octb200_params prm{};
	int N = 0, A = 0, B = 0, V = 1, H = 0, lines = 0, rawBytes = 2, R = 1;
	bool packed12 = false;            /* raw input is 12-bit packed (OCTB200_PACK_12P): 3 bytes per 2 samples */
	size_t inBytes = 0;               /* bytes of one raw input buffer */
	uint16_t* dUnpacked = nullptr;    /* u16 containers of a packed buffer, for the stages that cannot read it directly */
	long long S = 0;
	int device = 0, smCount = 148;
	int mode = OCTB200_FFT_FUSED;
	std::string err;

	cudaStream_t sCompute = nullptr, sH2D = nullptr, sD2H = nullptr;
	std::vector<void*> dRaw;
	std::vector<cudaEvent_t> evRawReady, evRawFree;
	int slot = -1;
	const void* lastDeviceRaw = nullptr;
	cudaEvent_t evTiming[8] = {};
	cudaEvent_t evComputeDone = nullptr, evFloatCopied = nullptr, evConvFree[2] = {};
	bool floatCopyPending = false, convPending[2] = { false, false };

	float* dVolumeOwned = nullptr;
	float* dVolume = nullptr;
	float* dTmp = nullptr;            /* S/2 floats: sinusoidal source */
	float2* dFft = nullptr;           /* S float2: SPLIT / CUFFT intermediate */
	float2* dFpnScratch = nullptr; size_t fpnScratchElems = 0;
	float2* dMeanLine = nullptr;
	float4* dFpnStats = nullptr; int fpnStatsBins = 0, fpnStatsSegW = 0;    /* [9][N]: candidates of the last determination (diagnostics) */
	float* dPpbg = nullptr;
	float* dPhase = nullptr; float2* dPhasor = nullptr;
	float4* dLutB = nullptr;       /* fused layout (de-interleaved by R) */
	float4* dLutB1 = nullptr;      /* natural order for the generic pre kernel */
	float2 *dTw = nullptr, *dCtw = nullptr;
	bool regKernel = false;           /* N in {1024, 2048} in a u16 container: the register kernels (k_fused.cuh) take the FUSED mode */
	bool genericOk = false;           /* the shared-memory kernel (k_generic.cu) can transform this line length */
	float2* dTwN = nullptr;           /* per-pass twiddle tables of the generic kernel */
	float4* dLutG = nullptr;          /* its natural-order 4-tap table (2 N entries) */
	int genRadix[16] = {}; int genPasses = 0; int genTwOff[16] = {}; unsigned genMagic[16] = {}; int genTwEntries = 1;
	float* dSinCurve = nullptr;
	void* dOutConv[2] = { nullptr, nullptr };
	int cufftPlan = -1;

	std::vector<float> hResample, hDispersion, hWindow, hPpbg;
	bool haveResample = false, haveDispersion = false, haveWindow = false;
	bool lutsDirty = true;
	int lutSa = -1, lutInterp = -1; bool lutRoll = false, lutWin = false, lutDisp = false;

	bool fpnDetermined = false;
	unsigned bufferNumberInVolume = 0, streamedBuffers = 0, streamingBufferNumber = 0, floatStreamingBufferNumber = 0, currentBufferNr = 0;
	void *hostBuf[2] = { nullptr, nullptr }; bool hostRegistered = false; bool hostBufMine[2] = { false, false };
	void *hostStream[2] = { nullptr, nullptr }; size_t hostStreamBytes = 0; bool hostStreamRegistered = false; bool hostStreamMine[2] = { false, false };
	void *hostFloat[2] = { nullptr, nullptr }; size_t hostFloatBytes = 0; bool hostFloatRegistered = false; bool hostFloatMine[2] = { false, false };
	octb200_host_callback cbStreaming = nullptr, cbFloat = nullptr, cbBackground = nullptr;
	unsigned long long launches = 0;
	unsigned long long pdlStamp = ~0ull;  /* value of `launches` right after the last main launch of the register kernel: the next main launch
	                                         may overlap its tail (programmatic dependent launch) only if nothing else was launched since */

	/* dispersion sweep scratch (grown on demand) */
	unsigned char* dSweepRaw = nullptr; size_t sweepRawBytes = 0;
	float4* dSweepLut = nullptr; size_t sweepLutElems = 0;
	float* dSweepOut = nullptr; size_t sweepOutElems = 0;
	float* dSweepPhase = nullptr; float2* dSweepPhasor = nullptr; size_t sweepPhaseElems = 0;
	float* dSweepMetric = nullptr; size_t sweepMetricElems = 0;

	/* en-face gather over peer memory (multi-GPU shards): local window = [1 KiB header: arrived[], ack[]][frame 0][frame 1][frame 2] (oct_device.cuh) */
	struct EnfaceGather {
		int world = 0, rank = 0;
		unsigned Eglobal = 0, offset = 0, seq = 0, consumedSeq = 0;
		unsigned char* window = nullptr;
		size_t frameStride = 0;
		unsigned char* peerBase[OCT_MAX_PEERS] = {};
		bool opened[OCT_MAX_PEERS] = {};
		bool connected = false;
		unsigned* counter = nullptr;       /* [0] producer kernels, [1] consume kernel, [2..3] status (time-outs: acks, arrivals) */
		float* display = nullptr;          /* private copy of the last consumed frame */
		bool autoOn = false;               /* every process call also gathers the en-face frame */
		unsigned autoFrame = 0, autoFrames = 1; int autoFn = 0;
	} eg;
};

namespace {

int fail(octb200_pipeline* p, int code, const char* fmt, ...) {
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
	if (p) p->err = buf; else g_createError = buf;
	return code;
}

#define CK(p, call)                                                                                   \
	do {                                                                                              \
		cudaError_t e_ = (call);                                                                      \
		if (e_ != cudaSuccess) return fail((p), (e_ == cudaErrorMemoryAllocation) ? OCTB200_ERR_NOMEM : OCTB200_ERR_CUDA, \
		                                   "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
	} while (0)

template <typename T> int dalloc(octb200_pipeline* p, T** ptr, size_t count) {
	CK(p, cudaMalloc((void**)ptr, count * sizeof(T)));
	CK(p, cudaMemset(*ptr, 0, count * sizeof(T)));     /* allocateAndInitializeBuffer, cuda_code.cu:975-1015 */
	return OCTB200_OK;
}
template <typename T> void dfree(T*& ptr) { if (ptr) { cudaFree(ptr); ptr = nullptr; } }
/* scratch that only ever grows */
template <typename T> int grow(octb200_pipeline* p, T*& ptr, size_t& have, size_t want) {
	if (have >= want) return OCTB200_OK;
	dfree(ptr); have = 0;
	int rc = dalloc(p, &ptr, want);
	if (rc == OCTB200_OK) have = want;
	return rc;
}

/* pin a caller-owned host buffer unless the caller already did (cudaHostAlloc / cudaHostRegister / torch pin_memory).
 * *mine tells whether we have to unpin it later. */
cudaError_t pin_host(void* h, size_t bytes, bool* mine) {
	*mine = false;
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost) return cudaSuccess;
	cudaGetLastError();
	cudaError_t e = cudaHostRegister(h, bytes, cudaHostRegisterPortable);
	if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return cudaSuccess; }
	if (e == cudaSuccess) *mine = true;
	return e;
}

void CUDART_CB host_cb_trampoline(void* data) {
	auto* pair = static_cast<std::pair<octb200_host_callback, void*>*>(data);
	if (pair->first) pair->first(pair->second);
	delete pair;
}
int enqueue_callback(octb200_pipeline* p, cudaStream_t st, octb200_host_callback cb, void* buf) {
	if (!cb) return OCTB200_OK;
	auto* pair = new std::pair<octb200_host_callback, void*>(cb, buf);
	CK(p, cudaLaunchHostFunc(st, host_cb_trampoline, pair));
	return OCTB200_OK;
}

/* stage selection from the parameter block == the dispatch table of cuda_code.cu:1448-1511 */
struct Stage { int sa; bool roll; int W; int HB, HA; };
Stage select_stage(const octb200_pipeline* p) {
	Stage s{};
	const octb200_params& q = p->prm;
	s.sa = q.resampling ? (q.resamplingInterpolation == OCTB200_INTERP_LANCZOS ? SA_LANCZOS : (q.resamplingInterpolation == OCTB200_INTERP_CUBIC ? SA_CUBIC : SA_LINEAR)) : SA_NONE;
	s.roll = q.backgroundRemoval != 0;
	s.W = q.rollingAverageWindowSize < 1 ? 1 : q.rollingAverageWindowSize;
	if (s.sa == SA_LANCZOS) { s.HB = 16 + (s.roll ? rup(s.W, 16) : 0); s.HA = s.HB; }
	else { s.HB = 0; s.HA = 0; }
	return s;
}

/* table flavour of the fused kernel's stage A (oct_luts.hpp build_stage_luts_paired): 4-tap stages carry two tap offsets, and the
 * R = 2 kernel without rolling mean reads a parity-split float slot */
int lut_taps_mode(int R, const Stage& st) {
	if (!(st.sa == SA_CUBIC || st.sa == SA_LINEAR)) return 0;
	return stage_a_splits_slot(R, st.sa, st.roll) ? 2 : 1;
}

int rebuild_luts(octb200_pipeline* p) {
	const octb200_params& q = p->prm;
	const Stage st = select_stage(p);
	const int N = p->N;
	if (q.resampling && !p->haveResample) return fail(p, OCTB200_ERR_NOT_READY, "resampling enabled but no resample curve set");
	if (q.dispersionCompensation && !p->haveDispersion) return fail(p, OCTB200_ERR_NOT_READY, "dispersion compensation enabled but no dispersion curve set");
	if (q.windowing && !p->haveWindow) return fail(p, OCTB200_ERR_NOT_READY, "windowing enabled but no window curve set");
	std::vector<float2> phasor;
	if (q.dispersionCompensation) {
		/* fillDispersivePhase on the device so that cos/sin are the same MUFU approximations the reference uses */
		CK(p, cudaMemcpyAsync(p->dPhase, p->hDispersion.data(), sizeof(float) * N, cudaMemcpyHostToDevice, p->sCompute));
		launch_fill_phase(p->dPhasor, p->dPhase, N, p->sCompute); p->launches++;
		phasor.resize(N);
		CK(p, cudaMemcpyAsync(phasor.data(), p->dPhasor, sizeof(float2) * N, cudaMemcpyDeviceToHost, p->sCompute));
		CK(p, cudaStreamSynchronize(p->sCompute));
	}
	const float* res = q.resampling ? p->hResample.data() : nullptr;
	const float* win = q.windowing ? p->hWindow.data() : nullptr;
	const float2* ph = q.dispersionCompensation ? phasor.data() : nullptr;
	const int interp = q.resamplingInterpolation;
	StageLuts l;
	if (p->N == 1024 || p->N == 2048) {
		std::vector<float4> paired;
		build_stage_luts_paired(N, p->R, interp == OCTB200_INTERP_CUBIC ? 1 : 0, res, win, ph, paired, lut_taps_mode(p->R, st));
		CK(p, cudaMemcpyAsync(p->dLutB, paired.data(), sizeof(float4) * 2 * N, cudaMemcpyHostToDevice, p->sCompute));
		CK(p, cudaStreamSynchronize(p->sCompute));
	}
	build_stage_luts(N, 1, res, win, ph, l);
	CK(p, cudaMemcpyAsync(p->dLutB1, l.B.data(), sizeof(float4) * N, cudaMemcpyHostToDevice, p->sCompute));
	CK(p, cudaStreamSynchronize(p->sCompute));
	if (p->dLutG) {
		std::vector<float4> taps;
		build_stage_luts_taps(N, interp == OCTB200_INTERP_CUBIC ? 1 : 0, res, win, ph, taps);
		CK(p, cudaMemcpyAsync(p->dLutG, taps.data(), sizeof(float4) * 2 * N, cudaMemcpyHostToDevice, p->sCompute));
		CK(p, cudaStreamSynchronize(p->sCompute));
	}
	p->lutsDirty = false;
	p->lutSa = st.sa; p->lutRoll = st.roll; p->lutInterp = interp; p->lutWin = q.windowing != 0; p->lutDisp = q.dispersionCompensation != 0;
	return OCTB200_OK;
}

EpiConsts epi_for(const octb200_pipeline* p, bool fpn, bool ppbg) {
	const octb200_params& q = p->prm;
	EpiConsts e = make_epi_consts(p->N, q.signalLogScaling, q.signalGrayscaleMin, q.signalGrayscaleMax, q.signalMultiplicator, q.signalAddend);
	e.fpn = fpn ? 1 : 0; e.ppbg = ppbg ? 1 : 0;
	e.ppbgWeight = q.postProcessBackgroundWeight; e.ppbgOffset = q.postProcessBackgroundOffset;
	return e;
}

int ensure_fft_buffer(octb200_pipeline* p) {
	if (!p->dFft) return dalloc(p, &p->dFft, (size_t)p->S);
	return OCTB200_OK;
}
int ensure_fpn_scratch(octb200_pipeline* p, size_t elems) {
	if (p->fpnScratchElems >= elems) return OCTB200_OK;
	dfree(p->dFpnScratch); p->fpnScratchElems = 0;
	int rc = dalloc(p, &p->dFpnScratch, elems);
	if (rc == OCTB200_OK) p->fpnScratchElems = elems;
	return rc;
}

/* cuda_bscanFlip runs over samplesPerBuffer/4 output elements (cuda_code.cu:794-805, launch :1547): with an odd number of B-scans
   per buffer the last (even-indexed) B-scan is only visited in its lower half and never swapped.  B-scan indices at or beyond this
   bound are not flipped; shards pass the B-scan count of the un-sharded buffer (octb200_config.bscansInUnshardedBuffer). */
unsigned flip_end(const octb200_pipeline* p) {
	const unsigned total = p->cfg.bscansInUnshardedBuffer ? p->cfg.bscansInUnshardedBuffer : p->cfg.bscanIndexBase + (unsigned)p->B;
	return total & ~1u;
}

FusedArgs fused_args(const octb200_pipeline* p, const Stage& st, const void* dRaw, int lines) {
	FusedArgs a{};
	a.raw = static_cast<const uint16_t*>(dRaw);
	a.cin = p->dFft;
	a.lutB = p->dLutB; a.tw = p->dTw; a.ctw = p->dCtw;
	a.meanLine = p->dMeanLine; a.ppbg = p->dPpbg;
	a.totalSamples = p->S; a.lines = lines; a.A = p->A;
	a.flip = p->prm.bscanFlip; a.bscanBase = p->cfg.bscanIndexBase; a.flipEnd = flip_end(p);
	a.shiftBits = p->prm.bitshift ? 4 : 0; a.W = st.W; a.HB = st.HB; a.HA = st.HA;
	a.lineBlock = 1;
	return a;
}
PreArgs pre_args(const octb200_pipeline* p, const Stage& st, const void* dRaw, int lines) {
	PreArgs a{};
	a.raw = dRaw; a.out = p->dFft; a.lutB = p->dLutB1;
	a.totalSamples = p->S; a.lines = lines; a.N = p->N;
	a.shiftBits = p->prm.bitshift ? 4 : 0; a.W = st.W;
	/* halos in elements, multiples of 16 so every container type keeps 16-byte alignment */
	a.HB = st.HB; a.HA = st.HA;
	const bool aligned = ((reinterpret_cast<uintptr_t>(dRaw) & 15) == 0) && (((size_t)p->N * p->rawBytes) % 16 == 0);
	a.useBulk = aligned ? 1 : 0;
	return a;
}

GenericArgs generic_args(const octb200_pipeline* p, const Stage& st, const void* dRaw, int lines) {
	GenericArgs a{};
	a.raw = dRaw; a.lutB = p->dLutB1; a.lutG = p->dLutG; a.tw = p->dTwN;
	a.meanLine = p->dMeanLine; a.ppbg = p->dPpbg;
	a.totalSamples = p->S; a.lines = lines; a.N = p->N; a.A = p->A;
	a.flip = p->prm.bscanFlip; a.bscanBase = p->cfg.bscanIndexBase; a.flipEnd = flip_end(p);
	a.shiftBits = p->prm.bitshift ? 4 : 0; a.W = st.W; a.HB = st.HB; a.HA = st.HA;
	const bool aligned = ((reinterpret_cast<uintptr_t>(dRaw) & 15) == 0) && (((size_t)p->N * p->rawBytes) % 16 == 0) && (((size_t)st.HB * p->rawBytes) % 16 == 0);
	a.useBulk = aligned ? 1 : 0;
	a.nPass = p->genPasses; a.twEntries = p->genTwEntries;
	for (int i = 0; i < 16; ++i) { a.radix[i] = p->genRadix[i]; a.twOff[i] = p->genTwOff[i]; a.magic[i] = p->genMagic[i]; }
	return a;
}

/* next en-face gather of this handle: sequence number, frame buffer seq % 3 in every rank's window, header words */
GatherDev next_gather(octb200_pipeline* p, unsigned frameNr, unsigned nFrames, int fn) {
	auto& g = p->eg;
	GatherDev d{};
	g.seq++;
	for (int r = 0; r < g.world; ++r) {
		d.frames[r] = reinterpret_cast<float*>(g.peerBase[r] + OCT_GATHER_HEADER_BYTES + (size_t)(g.seq % (unsigned)OCT_GATHER_FRAMES) * g.frameStride);
		d.flags[r] = reinterpret_cast<unsigned*>(g.peerBase[r]);
	}
	d.counter = g.counter; d.status = g.counter + 2; d.Eglobal = g.Eglobal; d.offset = g.offset; d.seq = g.seq;
	d.frameNr = (frameNr >= (unsigned)p->H) ? 0u : frameNr;          /* cuda_code.cu:1302 */
	d.nFrames = nFrames < 1 ? 1u : nFrames; d.fn = fn; d.world = g.world; d.rank = g.rank;
	return d;
}
cudaError_t launch_gather_standalone(octb200_pipeline* p, const GatherDev& d) {
	EnfaceGatherArgs a{};
	for (int r = 0; r < d.world; ++r) { a.frames[r] = d.frames[r]; a.flags[r] = d.flags[r]; }
	a.vol = p->dVolume; a.counter = d.counter; a.status = d.status;
	a.W = (unsigned)p->H; a.E = (unsigned)(p->A * p->B * p->V);
	a.frameNr = d.frameNr; a.nFrames = d.nFrames; a.fn = d.fn;
	a.Eglobal = d.Eglobal; a.offset = d.offset; a.world = d.world; a.rank = d.rank; a.seq = d.seq;
	return launch_enface_gather(a, p->sCompute);
}
/* consumer side of the latest gather (once per sequence number): wait for every rank's slab, copy the frame into the private display
   frame, acknowledge to every producer.  Enqueued behind every gather: a rank consumes each frame it takes part in, so the producers'
   flow-control wait (three gathers later) is satisfied long before they look. */
cudaError_t consume_gather(octb200_pipeline* p) {
	auto& g = p->eg;
	if (g.consumedSeq == g.seq) return cudaSuccess;
	EnfaceConsumeArgs a{};
	for (int r = 0; r < g.world; ++r) a.peerHeaders[r] = reinterpret_cast<unsigned*>(g.peerBase[r]);
	a.window = reinterpret_cast<const unsigned*>(g.window);
	a.frame = reinterpret_cast<const float*>(g.window + OCT_GATHER_HEADER_BYTES + (size_t)(g.seq % (unsigned)OCT_GATHER_FRAMES) * g.frameStride);
	a.display = g.display; a.counter = g.counter + 1; a.status = g.counter + 2;
	a.Eglobal = g.Eglobal; a.seq = g.seq; a.world = g.world; a.rank = g.rank;
	/* in stream order, as the programmatic dependent of the kernel that produced this rank's slab.  (A consumer kernel on a stream of its
	   own was measured and dropped: a saturated compute stream starves it -- every SM goes to the next, already queued compute grid --
	   and a compute grid that spins in its prologue for acknowledgements then holds every SM the consumer needs: 10 s time-outs per
	   buffer at 8 GPUs, profiles/r02l_bench_n8_display_stream_stall.json.  In order it can never be starved, and the acknowledgement of frame s is on its
	   way before the kernel of s + 1 starts, three buffers ahead of need.) */
	const bool dependent = (p->pdlStamp == p->launches) && !(p->cfg.flags & OCTB200_FLAG_NO_DEPENDENT_LAUNCH);
	cudaError_t e = launch_enface_consume(a, p->smCount, dependent, p->sCompute);
	if (e == cudaSuccess) {
		g.consumedSeq = g.seq;
		const bool chain = p->pdlStamp == p->launches;
		p->launches++;
		if (chain) p->pdlStamp = p->launches;      /* the next main launch may be this kernel's programmatic dependent in turn */
	}
	return e;
}

/* a line group of the fused kernel works through blocks of this many consecutive lines when the en-face gather is fused into its
   epilogue (one coalesced peer store per block and rank): the largest of 8, 4, 2, 1 that does not add a round to the slowest group */
int gather_line_block(const octb200_pipeline* p, const Stage& st, int src, int lines) {
	int grid = 0, threads = 0, smem = 0;
	fused_launch_shape(p->R, st.sa, st.roll, src, st.HB, st.HA, p->smCount, lines, &grid, &threads, &smem);
	const int G = grid * (threads / 32 / p->R);
	if (G < 1) return 1;
	const int rounds = (lines + G - 1) / G;
	for (int lb = 8; lb > 1; lb >>= 1) {
		const int blocks = (lines + lb - 1) / lb;
		if (((blocks + G - 1) / G) * lb == rounds) return lb;
	}
	return 1;
}

/* the whole per-buffer chain on the compute stream; dRaw is device memory */
int run_chain(octb200_pipeline* p, const void* dRaw) {
	octb200_params& q = p->prm;
	const Stage st = select_stage(p);
	if (p->lutsDirty || st.sa != p->lutSa || st.roll != p->lutRoll) { int rc = rebuild_luts(p); if (rc) return rc; }
	if (q.postProcessBackgroundRemoval && p->hPpbg.empty() && !q.postProcessBackgroundRecordingRequested) {
		/* the reference starts with a zeroed background line (cuda_code.cu:1122) */
	}

	/* slab of the processed volume (cuda_code.cu:1530-1535) */
	if (p->V > 1) p->bufferNumberInVolume = (p->bufferNumberInVolume + 1) % (unsigned)p->V;
	float* slab = p->dVolume + (size_t)(p->S / 2) * p->bufferNumberInVolume;

	int mode = p->mode;
	bool generic = false;              /* FUSED by the shared-memory kernel (line lengths / containers the register kernels do not take) */
	/* where the register kernel reads its lines from: u16 containers, 12-bit packed, or u8 / u32 containers */
	int rawSrc = p->rawBytes == 1 ? SRC_RAW8 : (p->rawBytes == 4 ? SRC_RAW32 : SRC_RAW16);
	if (mode == OCTB200_FFT_FUSED && p->regKernel) {
		/* a huge rolling window with Lanczos halos may not fit the fused kernel's shared memory; the u8 / u32 slot conversions exist for
		   the 4-tap and plain stages without rolling mean -- everything else takes the split chain (pre kernel + register FFT kernel) */
		int g = 0, t = 0, sm = 0;
		fused_launch_shape(p->R, st.sa, st.roll, SRC_RAW16, st.HB, st.HA, p->smCount, p->lines, &g, &t, &sm);
		if (t < 32 * p->R || (rawSrc != SRC_RAW16 && (st.roll || st.sa == SA_LANCZOS))) mode = OCTB200_FFT_SPLIT;
	} else if (mode == OCTB200_FFT_FUSED) {
		generic = generic_fits(p->N, p->rawBytes, st.HB, st.HA, st.roll, p->genTwEntries);
		if (!generic) mode = OCTB200_FFT_CUFFT;
	}
	if (mode != OCTB200_FFT_FUSED) { int rc = ensure_fft_buffer(p); if (rc) return rc; }

	/* 12-bit packed input: the fused kernel unpacks it in its slot conversion (4-tap / plain stage, no rolling mean); every other
	   stage reads u16 containers, so the buffer is unpacked once into HBM first */
	if (p->packed12) {
		if (mode == OCTB200_FFT_FUSED && st.sa != SA_LANCZOS && !st.roll) rawSrc = SRC_RAW12P;
		else {
			if (!p->dUnpacked) { int rc = dalloc(p, &p->dUnpacked, (size_t)p->S + 32); if (rc) return rc; }
			CK(p, launch_unpack12(p->dUnpacked, dRaw, p->S / 8, p->smCount, p->sCompute)); p->launches++;
			dRaw = p->dUnpacked;
		}
	}

	if (p->floatCopyPending) { CK(p, cudaStreamWaitEvent(p->sCompute, p->evFloatCopied, 0)); p->floatCopyPending = false; }

	const bool fpn = q.fixedPatternNoiseRemoval != 0;
	const bool sinus = q.sinusoidalScanCorrection != 0;
	const bool ppbgOn = q.postProcessBackgroundRemoval != 0;
	const bool ppbgRecord = ppbgOn && q.postProcessBackgroundRecordingRequested;
	const bool ppbgFoldMain = ppbgOn && !ppbgRecord && !sinus;
	const bool ppbgFoldSinus = ppbgOn && !ppbgRecord && sinus;
	float* mainOut = slab;
	if (sinus) { if (!p->dTmp) { int rc = dalloc(p, &p->dTmp, (size_t)(p->S / 2)); if (rc) return rc; } mainOut = p->dTmp; }

	const bool determine = fpn && ((!q.continuousFixedPatternNoiseDetermination && !p->fpnDetermined) ||
	                               q.continuousFixedPatternNoiseDetermination || q.redetermineFixedPatternNoise);   /* cuda_code.cu:1521 */
	int fpnHeight = (int)q.bscansForNoiseDetermination * p->A;
	if (fpnHeight > p->lines) fpnHeight = p->lines;

	/* converted output for the stream-to-host path (cuda_code.cu:1357-1372): which of the two device buffers this call fills, and
	   whether floatToOutput is folded into the main kernel's epilogue (the slab must be final after it, u16 containers) */
	const bool wantConv = q.streamToHost && p->hostStream[0] && p->hostStream[1];
	const bool convThisBuffer = wantConv && (p->streamedBuffers % (q.streamingBuffersToSkip + 1) == 0);     /* cuda_code.cu:1358 */
	const int convSlot = convThisBuffer ? (int)((p->streamingBufferNumber + 1) % 2) : -1;
	const bool convFoldable = convThisBuffer && !ppbgRecord && p->rawBytes == 2 && !(p->cfg.flags & OCTB200_FLAG_SEPARATE_CONVERSION);
	const bool convInSinus = convFoldable && sinus;                                  /* the sinusoidal kernel writes the final slab: any FFT mode */
	const bool convFused = convFoldable && !sinus && mode == OCTB200_FFT_FUSED && !generic;      /* the register kernel's epilogue does */
	const float convScale = (float)((1u << (p->cfg.bitDepth <= 10 ? 10 : p->cfg.bitDepth <= 12 ? 12 : 16)) - 1u);   /* cuda_code.cu:948-958 */

	bool gatherDone = false;
	if (mode == OCTB200_FFT_CUFFT) {
		PreArgs pa = pre_args(p, st, dRaw, p->lines);
		CK(p, launch_pre(pa, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute)); p->launches++;
		if (p->cufftPlan < 0) {
			if (!g_cufft.load(p->err)) return OCTB200_ERR_CUDA;
			int plan = 0;
			if (g_cufft.Plan1d(&plan, p->N, kCufftC2C, p->lines) != 0) return fail(p, OCTB200_ERR_CUDA, "cufftPlan1d failed");
			p->cufftPlan = plan;
			g_cufft.SetStream(plan, p->sCompute);
		}
		if (g_cufft.ExecC2C(p->cufftPlan, p->dFft, p->dFft, kCufftInverse) != 0) return fail(p, OCTB200_ERR_CUDA, "cufftExecC2C failed");
		p->launches++;
		if (determine) {
			CK(p, launch_fpn_minvar(p->dMeanLine, p->dFft, p->N, p->N, fpnHeight, p->dFpnStats, p->sCompute)); p->launches++;
			p->fpnStatsBins = p->N; p->fpnStatsSegW = fpnHeight / 9;
			p->fpnDetermined = true; q.redetermineFixedPatternNoise = 0;
		}
		PostArgs po{};
		po.in = p->dFft; po.out = mainOut; po.meanLine = p->dMeanLine; po.ppbg = p->dPpbg;
		po.epi = epi_for(p, fpn, ppbgFoldMain); po.lines = p->lines; po.N = p->N; po.A = p->A;
		po.flip = q.bscanFlip; po.bscanBase = p->cfg.bscanIndexBase; po.flipEnd = flip_end(p);
		CK(p, launch_post(po, p->smCount, p->sCompute)); p->launches++;
	} else if (generic) {
		if (determine) {
			int rc = ensure_fpn_scratch(p, (size_t)fpnHeight * p->H); if (rc) return rc;
			GenericArgs ga = generic_args(p, st, dRaw, fpnHeight);
			ga.cplxOut = p->dFpnScratch; ga.epi = epi_for(p, false, false);
			CK(p, launch_generic(ga, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute)); p->launches++;
			CK(p, launch_fpn_minvar(p->dMeanLine, p->dFpnScratch, p->H, p->H, fpnHeight, p->dFpnStats, p->sCompute)); p->launches++;
			p->fpnStatsBins = p->H; p->fpnStatsSegW = fpnHeight / 9;
			p->fpnDetermined = true; q.redetermineFixedPatternNoise = 0;
		}
		GenericArgs ga = generic_args(p, st, dRaw, p->lines);
		ga.out = mainOut; ga.epi = epi_for(p, fpn, ppbgFoldMain);
		CK(p, launch_generic(ga, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute)); p->launches++;
	} else {
		const int src = (mode == OCTB200_FFT_FUSED) ? rawSrc : SRC_CPLX;
		if (determine) {
			int rc = ensure_fpn_scratch(p, (size_t)fpnHeight * p->H); if (rc) return rc;
			if (src == SRC_CPLX) { PreArgs pa = pre_args(p, st, dRaw, fpnHeight); CK(p, launch_pre(pa, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute)); p->launches++; }
			FusedArgs fa = fused_args(p, st, dRaw, fpnHeight);
			fa.cplxOut = p->dFpnScratch; fa.epi = epi_for(p, false, false);
			CK(p, launch_fused(p->R, st.sa, st.roll, src, fa, p->smCount, p->sCompute)); p->launches++;
			CK(p, launch_fpn_minvar(p->dMeanLine, p->dFpnScratch, p->H, p->H, fpnHeight, p->dFpnStats, p->sCompute)); p->launches++;
			p->fpnStatsBins = p->H; p->fpnStatsSegW = fpnHeight / 9;
			p->fpnDetermined = true; q.redetermineFixedPatternNoise = 0;
		}
		if (src == SRC_CPLX) { PreArgs pa = pre_args(p, st, dRaw, p->lines); CK(p, launch_pre(pa, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute)); p->launches++; }
		FusedArgs fa = fused_args(p, st, dRaw, p->lines);
		fa.out = mainOut; fa.epi = epi_for(p, fpn, ppbgFoldMain);
		/* automatic en-face gather: fused into this kernel's epilogue when the slab is final after it */
		if (p->eg.autoOn && p->eg.connected && !sinus && !ppbgRecord && p->V == 1 && p->eg.autoFrames <= 1) {
			fa.eg = next_gather(p, p->eg.autoFrame, p->eg.autoFrames, p->eg.autoFn);
			fa.lineBlock = gather_line_block(p, st, src, p->lines);
			gatherDone = true;
		}
		if (convFused) {
			if (p->convPending[convSlot]) { CK(p, cudaStreamWaitEvent(p->sCompute, p->evConvFree[convSlot], 0)); p->convPending[convSlot] = false; }
			fa.convOut = static_cast<unsigned short*>(p->dOutConv[convSlot]);
			fa.convScale = convScale;
		}
		/* back-to-back buffers: when the previous launch of this handle was this same main kernel (nothing in between that writes a table
		   the prologue reads), let this grid's prologue overlap its tail */
		fa.pdl = (src != SRC_CPLX && !determine && p->pdlStamp == p->launches && !(p->cfg.flags & OCTB200_FLAG_NO_DEPENDENT_LAUNCH)) ? 1 : 0;
		CK(p, launch_fused(p->R, st.sa, st.roll, src, fa, p->smCount, p->sCompute)); p->launches++;
		p->pdlStamp = p->launches;
	}

	if (sinus) {
		if (convInSinus && p->convPending[convSlot]) { CK(p, cudaStreamWaitEvent(p->sCompute, p->evConvFree[convSlot], 0)); p->convPending[convSlot] = false; }
		CK(p, launch_sinusoidal(slab, p->dTmp, p->dSinCurve, p->H, p->A, p->S / 2, ppbgFoldSinus ? 1 : 0, p->dPpbg,
		                        q.postProcessBackgroundWeight, q.postProcessBackgroundOffset,
		                        convInSinus ? static_cast<unsigned short*>(p->dOutConv[convSlot]) : nullptr, convScale, p->smCount, p->sCompute)); p->launches++;
	}
	if (ppbgRecord) {
		/* cuda_code.cu:1558-1567: record from the first B-scan of this buffer, hand it to the host, then remove */
		CK(p, launch_ppbg_record(p->dPpbg, slab, p->H, p->A, p->sCompute)); p->launches++;
		p->hPpbg.resize(p->H);
		CK(p, cudaMemcpyAsync(p->hPpbg.data(), p->dPpbg, sizeof(float) * p->H, cudaMemcpyDeviceToHost, p->sCompute));
		int rc = enqueue_callback(p, p->sCompute, p->cbBackground, p->hPpbg.data()); if (rc) return rc;
		q.postProcessBackgroundRecordingRequested = 0;
		CK(p, launch_ppbg_remove(slab, p->dPpbg, q.postProcessBackgroundWeight, q.postProcessBackgroundOffset, p->H, p->S / 2, p->smCount, p->sCompute)); p->launches++;
	}

	if (p->eg.autoOn && p->eg.connected && !gatherDone) {
		/* later passes changed the slab (or the volume has several slabs): stand-alone extraction + peer stores */
		const GatherDev d = next_gather(p, p->eg.autoFrame, p->eg.autoFrames, p->eg.autoFn);
		CK(p, launch_gather_standalone(p, d)); p->launches++;
	}
	if (p->eg.autoOn && p->eg.connected) CK(p, consume_gather(p));

	/* ---- streaming to the host (cuda_code.cu:1357-1386,1595-1604) ---- */
	const bool wantFloat = q.streamFloatToHost && p->hostFloat[0] && p->hostFloat[1];
	if (wantFloat || wantConv) CK(p, cudaEventRecord(p->evComputeDone, p->sCompute));
	if (wantFloat) {
		p->floatStreamingBufferNumber = (p->floatStreamingBufferNumber + 1) % 2;
		void* dst = p->hostFloat[p->floatStreamingBufferNumber];
		CK(p, cudaStreamWaitEvent(p->sD2H, p->evComputeDone, 0));
		CK(p, cudaMemcpyAsync(dst, slab, (size_t)(p->S / 2) * sizeof(float), cudaMemcpyDeviceToHost, p->sD2H));
		CK(p, cudaEventRecord(p->evFloatCopied, p->sD2H)); p->floatCopyPending = true;
		int rc = enqueue_callback(p, p->sD2H, p->cbFloat, dst); if (rc) return rc;
	}
	if (wantConv) {
		p->currentBufferNr = p->bufferNumberInVolume;
		if (convThisBuffer) {
			p->streamedBuffers = 0;
			p->streamingBufferNumber = (unsigned)convSlot;
			const int i = convSlot;
			if (!convFused && !convInSinus) {
				if (p->convPending[i]) { CK(p, cudaStreamWaitEvent(p->sCompute, p->evConvFree[i], 0)); p->convPending[i] = false; }
				CK(p, launch_float_to_output(p->dOutConv[i], slab, (int)p->cfg.bitDepth, p->S / 2, p->smCount, p->sCompute)); p->launches++;
				CK(p, cudaEventRecord(p->evComputeDone, p->sCompute));
			}
			CK(p, cudaStreamWaitEvent(p->sD2H, p->evComputeDone, 0));
			CK(p, cudaMemcpyAsync(p->hostStream[i], p->dOutConv[i], (size_t)(p->S / 2) * p->rawBytes, cudaMemcpyDeviceToHost, p->sD2H));
			CK(p, cudaEventRecord(p->evConvFree[i], p->sD2H)); p->convPending[i] = true;
			int rc = enqueue_callback(p, p->sD2H, p->cbStreaming, p->hostStream[i]); if (rc) return rc;
		}
		p->streamedBuffers++;
	}
	return OCTB200_OK;
}

int use_device(const octb200_pipeline* p) { return cudaSetDevice(p->device) == cudaSuccess ? 0 : -1; }

}  // namespace

/* ====================================================================== C ABI */

extern "C" {

int octb200_version(void) { return OCTB200_VERSION; }

void octb200_default_params(octb200_params* o) {
	if (!o) return;
	std::memset(o, 0, sizeof(*o));
	o->signalGrayscaleMin = 0.0f; o->signalGrayscaleMax = 60.0f; o->signalMultiplicator = 1.0f; o->signalAddend = 0.0f;
	o->rollingAverageWindowSize = 1; o->bscansForNoiseDetermination = 1;
	o->postProcessBackgroundWeight = 1.0f; o->postProcessBackgroundOffset = 0.0f;
}

const char* octb200_last_error(const octb200_pipeline* p) { return p ? p->err.c_str() : g_createError.c_str(); }
int octb200_effective_fft_mode(const octb200_pipeline* p) { return p ? p->mode : OCTB200_ERR_INVALID; }

int octb200_query_fft_path(uint32_t samplesPerLine, uint32_t bitDepth, int32_t* radices, int32_t* nPasses) {
	if (samplesPerLine < 8 || (samplesPerLine & 1) || bitDepth < 1 || bitDepth > 32) return OCTB200_ERR_INVALID;
	const int N = (int)samplesPerLine;
	const int rawBytes = bitDepth <= 8 ? 1 : (bitDepth <= 16 ? 2 : 4);
	if (nPasses) *nPasses = 0;
	if (N == 1024 || N == 2048) return OCTB200_PATH_REGISTER_KERNEL;
	int radix[16] = {}, passes = 0, twOff[16]; unsigned magic[16];
	bool generic = generic_fft_plan(N, radix, &passes);
	if (generic) generic = generic_fits(N, rawBytes, 0, 0, false, generic_twiddle_layout(N, radix, passes, twOff, magic));
	if (generic) {
		if (nPasses) *nPasses = passes;
		if (radices) for (int i = 0; i < passes; ++i) radices[i] = radix[i];
		const bool pow2 = (N & (N - 1)) == 0;
		return (!pow2 || N > 2048) ? OCTB200_PATH_SHARED_MEMORY_KERNEL : OCTB200_PATH_CUFFT_CHAIN_SHARED_AVAILABLE;
	}
	return OCTB200_PATH_CUFFT_CHAIN;
}

int octb200_create(const octb200_config* cfg, octb200_pipeline** out) {
	if (!cfg || !out) return fail(nullptr, OCTB200_ERR_INVALID, "null argument");
	*out = nullptr;
	if (cfg->samplesPerLine < 8 || (cfg->samplesPerLine & 1) || cfg->ascansPerBscan < 1 || cfg->bscansPerBuffer < 1 ||
	    cfg->buffersPerVolume < 1 || cfg->bitDepth < 1 || cfg->bitDepth > 32)
		return fail(nullptr, OCTB200_ERR_INVALID, "invalid acquisition geometry");
	const long long S = (long long)cfg->samplesPerLine * cfg->ascansPerBscan * cfg->bscansPerBuffer;
	if (S >= (1LL << 31)) return fail(nullptr, OCTB200_ERR_INVALID, "samples per buffer must be < 2^31 (as in the reference)");
	int dev = cfg->device;
	if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, OCTB200_ERR_CUDA, "no CUDA device"); }
	if (cudaSetDevice(dev) != cudaSuccess) return fail(nullptr, OCTB200_ERR_CUDA, "cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(cudaGetLastError()));

	auto* p = new octb200_pipeline();
	p->cfg = *cfg; p->device = dev;
	octb200_default_params(&p->prm);
	p->N = (int)cfg->samplesPerLine; p->A = (int)cfg->ascansPerBscan; p->B = (int)cfg->bscansPerBuffer; p->V = (int)cfg->buffersPerVolume;
	p->H = p->N / 2; p->lines = p->A * p->B; p->S = S;
	p->rawBytes = cfg->bitDepth <= 8 ? 1 : (cfg->bitDepth <= 16 ? 2 : 4);   /* bytesPerSample, cuda_code.cu:1077 */
	p->R = p->N == 2048 ? 2 : 1;
	p->packed12 = cfg->inputPacking == OCTB200_PACK_12P;
	if (cfg->inputPacking > OCTB200_PACK_12P || (p->packed12 && (cfg->bitDepth != 12 || (cfg->samplesPerLine % 32) != 0))) {
		delete p;
		return fail(nullptr, OCTB200_ERR_INVALID, "inputPacking %u needs bitDepth 12 and samplesPerLine a multiple of 32", cfg->inputPacking);
	}
	p->inBytes = p->packed12 ? (size_t)S * 3 / 2 : (size_t)S * p->rawBytes;
	const bool fftSize = (p->N == 1024 || p->N == 2048);
	p->regKernel = fftSize;             /* u16 directly; u8 / u32 containers through the SRC_RAW8 / SRC_RAW32 slot conversions (4-tap and plain stages) */
	p->genericOk = !p->packed12 && generic_fft_plan(p->N, p->genRadix, &p->genPasses);
	if (p->genericOk) {
		p->genTwEntries = generic_twiddle_layout(p->N, p->genRadix, p->genPasses, p->genTwOff, p->genMagic);
		p->genericOk = generic_fits(p->N, p->rawBytes, 0, 0, false, p->genTwEntries);
	}
	int mode = cfg->fftMode;
	/* FUSED = one kernel from raw samples to B-scan lines: the register kernels where they apply, else the shared-memory kernel */
	/* AUTO: the register kernels wherever they apply; the shared-memory kernel where it beats the cuFFT chain on a B200 (measured,
	   profiles/r02g_generic_and_container_perf.json: line lengths that are not a power of two, and N > 2048; for short power-of-two lines the three
	   kernel chain around cuFFT is as fast or faster); the cuFFT chain otherwise.  An explicit FUSED is honoured wherever a fused kernel exists. */
	const bool pow2 = (p->N & (p->N - 1)) == 0;
	if (mode == OCTB200_FFT_AUTO) mode = p->regKernel ? OCTB200_FFT_FUSED : ((p->genericOk && (!pow2 || p->N > 2048)) ? OCTB200_FFT_FUSED : OCTB200_FFT_CUFFT);
	if ((mode == OCTB200_FFT_FUSED && !(p->regKernel || p->genericOk)) || (mode == OCTB200_FFT_SPLIT && !fftSize) ||
	    mode < OCTB200_FFT_FUSED || mode > OCTB200_FFT_CUFFT) {
		delete p;
		return fail(nullptr, OCTB200_ERR_INVALID, "fftMode %d unsupported for samplesPerLine=%u bitDepth=%u (FUSED: even N <= 8192 with prime factors <= 13; "
		            "SPLIT: N in {1024,2048}; packed 12-bit input: N in {1024,2048})", cfg->fftMode, cfg->samplesPerLine, cfg->bitDepth);
	}
	p->mode = mode;

	int rc = OCTB200_OK;
	auto bail = [&](int code) { g_createError = p->err; octb200_destroy(p); return code; };
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return bail(fail(p, OCTB200_ERR_CUDA, "cudaGetDeviceProperties failed"));
	p->smCount = prop.multiProcessorCount;
	if (prop.major < 10) return bail(fail(p, OCTB200_ERR_INVALID, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major, prop.minor));

#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(p, e_ == cudaErrorMemoryAllocation ? OCTB200_ERR_NOMEM : OCTB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
#define RCC(call) do { rc = (call); if (rc) return bail(rc); } while (0)
	CKC(cudaStreamCreateWithFlags(&p->sCompute, cudaStreamNonBlocking));
	CKC(cudaStreamCreateWithFlags(&p->sH2D, cudaStreamNonBlocking));
	CKC(cudaStreamCreateWithFlags(&p->sD2H, cudaStreamNonBlocking));
	for (auto& e : p->evTiming) CKC(cudaEventCreate(&e));
	CKC(cudaEventCreateWithFlags(&p->evComputeDone, cudaEventDisableTiming));
	CKC(cudaEventCreateWithFlags(&p->evFloatCopied, cudaEventDisableTiming));
	for (auto& e : p->evConvFree) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

	const int slots = cfg->rawSlots > 0 ? cfg->rawSlots : 2;
	p->dRaw.assign(slots, nullptr); p->evRawReady.assign(slots, nullptr); p->evRawFree.assign(slots, nullptr);
	for (int i = 0; i < slots; ++i) {
		unsigned char* r = nullptr;
		RCC(dalloc(p, &r, p->inBytes + 64)); p->dRaw[i] = r;
		CKC(cudaEventCreateWithFlags(&p->evRawReady[i], cudaEventDisableTiming));
		CKC(cudaEventCreateWithFlags(&p->evRawFree[i], cudaEventDisableTiming));
	}
	RCC(dalloc(p, &p->dVolumeOwned, (size_t)(S / 2) * p->V)); p->dVolume = p->dVolumeOwned;
	RCC(dalloc(p, &p->dMeanLine, (size_t)p->N));
	RCC(dalloc(p, &p->dFpnStats, (size_t)9 * p->N));
	RCC(dalloc(p, &p->dPpbg, (size_t)p->H));
	RCC(dalloc(p, &p->dPhase, (size_t)p->N));
	RCC(dalloc(p, &p->dPhasor, (size_t)p->N));
	RCC(dalloc(p, &p->dLutB, (size_t)2 * p->N)); RCC(dalloc(p, &p->dLutB1, (size_t)p->N));
	RCC(dalloc(p, &p->dTw, (size_t)1024)); RCC(dalloc(p, &p->dCtw, (size_t)1024));
	if (p->genericOk && !p->regKernel) {
		const int entries = p->genTwEntries;
		RCC(dalloc(p, &p->dTwN, (size_t)entries));
		RCC(dalloc(p, &p->dLutG, (size_t)2 * p->N));
		std::vector<float2> twn((size_t)entries, make_float2(1.f, 0.f));
		generic_fill_twiddles(p->genRadix, p->genPasses, p->genTwOff, twn.data());
		CKC(cudaMemcpy(p->dTwN, twn.data(), sizeof(float2) * entries, cudaMemcpyHostToDevice));
	}
	RCC(dalloc(p, &p->dSinCurve, (size_t)p->A));
	{
		unsigned char* c0 = nullptr; unsigned char* c1 = nullptr;
		RCC(dalloc(p, &c0, (size_t)(S / 2) * p->rawBytes)); p->dOutConv[0] = c0;
		RCC(dalloc(p, &c1, (size_t)(S / 2) * p->rawBytes)); p->dOutConv[1] = c1;
	}
	{
		std::vector<float2> tw, ctw; build_twiddles_1024(tw); build_combine_twiddles_2048(ctw);
		CKC(cudaMemcpy(p->dTw, tw.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice));
		CKC(cudaMemcpy(p->dCtw, ctw.data(), sizeof(float2) * 1024, cudaMemcpyHostToDevice));
		std::vector<float> sc(p->A); curves::sinusoidal(p->A, sc.data());    /* cuda_code.cu:1093 */
		CKC(cudaMemcpy(p->dSinCurve, sc.data(), sizeof(float) * p->A, cudaMemcpyHostToDevice));
	}
#undef CKC
#undef RCC
	p->bufferNumberInVolume = (unsigned)p->V - 1;   /* cuda_code.cu:1146 */
	*out = p;
	return OCTB200_OK;
}

int octb200_destroy(octb200_pipeline* p) {
	if (!p) return OCTB200_OK;
	cudaSetDevice(p->device);
	cudaDeviceSynchronize();
	if (p->hostRegistered) octb200_unregister_host_buffers(p);
	if (p->hostStreamRegistered) octb200_unregister_streaming_buffers(p);
	if (p->hostFloatRegistered) octb200_unregister_float_streaming_buffers(p);
	if (p->cufftPlan >= 0 && g_cufft.Destroy) g_cufft.Destroy(p->cufftPlan);
	for (void*& r : p->dRaw) { if (r) cudaFree(r); r = nullptr; }
	for (auto e : p->evRawReady) if (e) cudaEventDestroy(e);
	for (auto e : p->evRawFree) if (e) cudaEventDestroy(e);
	for (auto e : p->evTiming) if (e) cudaEventDestroy(e);
	if (p->evComputeDone) cudaEventDestroy(p->evComputeDone);
	if (p->evFloatCopied) cudaEventDestroy(p->evFloatCopied);
	for (auto e : p->evConvFree) if (e) cudaEventDestroy(e);
	dfree(p->dVolumeOwned); dfree(p->dTmp); dfree(p->dFft); dfree(p->dFpnScratch); dfree(p->dMeanLine); dfree(p->dFpnStats); dfree(p->dPpbg);
	dfree(p->dPhase); dfree(p->dPhasor); dfree(p->dLutB); dfree(p->dLutB1);
	dfree(p->dTw); dfree(p->dCtw); dfree(p->dTwN); dfree(p->dLutG); dfree(p->dSinCurve);
	for (void*& c : p->dOutConv) { if (c) cudaFree(c); c = nullptr; }
	octb200_enface_gather_close(p);
	dfree(p->dUnpacked);
	dfree(p->dSweepRaw); dfree(p->dSweepLut); dfree(p->dSweepOut); dfree(p->dSweepPhase); dfree(p->dSweepPhasor); dfree(p->dSweepMetric);
	if (p->sCompute) cudaStreamDestroy(p->sCompute);
	if (p->sH2D) cudaStreamDestroy(p->sH2D);
	if (p->sD2H) cudaStreamDestroy(p->sD2H);
	delete p;
	return OCTB200_OK;
}

/* ---------- parameters and curves ---------- */
int octb200_set_params(octb200_pipeline* p, const octb200_params* prm) {
	if (!p || !prm) return fail(p, OCTB200_ERR_INVALID, "null argument");
	const octb200_params old = p->prm;
	p->prm = *prm;
	if (old.resampling != prm->resampling || old.resamplingInterpolation != prm->resamplingInterpolation ||
	    old.windowing != prm->windowing || old.dispersionCompensation != prm->dispersionCompensation)
		p->lutsDirty = true;
	return OCTB200_OK;
}

static int set_curve(octb200_pipeline* p, std::vector<float>& dst, bool& have, const float* src, int n, const char* what) {
	if (!p || !src) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (n != p->N) return fail(p, OCTB200_ERR_INVALID, "%s curve length %d != samplesPerLine %d", what, n, p->N);
	dst.assign(src, src + n); have = true; p->lutsDirty = true;
	return OCTB200_OK;
}
int octb200_set_resample_curve(octb200_pipeline* p, const float* c, int n) {
	int rc = set_curve(p, p->hResample, p->haveResample, c, n, "resample");
	if (rc == OCTB200_OK) curves::clamp_resample(n, p->hResample.data());   /* octalgorithmparameters.cpp:167 also clamps custom curves */
	return rc;
}
int octb200_set_dispersion_curve(octb200_pipeline* p, const float* c, int n) { return set_curve(p, p->hDispersion, p->haveDispersion, c, n, "dispersion"); }
int octb200_set_window_curve(octb200_pipeline* p, const float* c, int n) { return set_curve(p, p->hWindow, p->haveWindow, c, n, "window"); }

int octb200_set_postprocess_background(octb200_pipeline* p, const float* bg, int n) {
	if (!p || !bg) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (n != p->H) return fail(p, OCTB200_ERR_INVALID, "background length %d != samplesPerLine/2 %d", n, p->H);
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	p->hPpbg.assign(bg, bg + n);
	CK(p, cudaMemcpyAsync(p->dPpbg, p->hPpbg.data(), sizeof(float) * n, cudaMemcpyHostToDevice, p->sCompute));
	CK(p, cudaStreamSynchronize(p->sCompute));
	return OCTB200_OK;
}
int octb200_get_postprocess_background(octb200_pipeline* p, float* bg, int n) {
	if (!p || !bg || n != p->H) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaStreamSynchronize(p->sCompute));
	CK(p, cudaMemcpy(bg, p->dPpbg, sizeof(float) * n, cudaMemcpyDeviceToHost));
	return OCTB200_OK;
}
int octb200_get_fpn_mean_line(octb200_pipeline* p, float* reIm, int n) {
	if (!p || !reIm || n != p->N) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaStreamSynchronize(p->sCompute));
	CK(p, cudaMemcpy(reIm, p->dMeanLine, sizeof(float2) * n, cudaMemcpyDeviceToHost));
	return OCTB200_OK;
}
int octb200_get_fpn_segment_stats(octb200_pipeline* p, float* stats, int bins, int* segmentLength) {
	if (!p || !stats || bins < 1 || bins > p->H) return fail(p, OCTB200_ERR_INVALID, "bad argument (bins must be 1 .. samplesPerLine/2)");
	if (p->fpnStatsBins < 1) return fail(p, OCTB200_ERR_NOT_READY, "no fixed-pattern-noise determination has run on this handle yet");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaStreamSynchronize(p->sCompute));
	/* device layout [9][fpnStatsBins] (H bins on the own-FFT paths, N on the cuFFT path) -> caller's [9][bins] */
	CK(p, cudaMemcpy2D(stats, (size_t)bins * sizeof(float4), p->dFpnStats, (size_t)p->fpnStatsBins * sizeof(float4), (size_t)bins * sizeof(float4), 9,
	                   cudaMemcpyDeviceToHost));
	if (segmentLength) *segmentLength = p->fpnStatsSegW;
	return OCTB200_OK;
}
int octb200_set_fpn_mean_line(octb200_pipeline* p, const float* reIm, int n) {
	if (!p || !reIm || n != p->N) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaStreamSynchronize(p->sCompute));
	CK(p, cudaMemcpy(p->dMeanLine, reIm, sizeof(float2) * n, cudaMemcpyHostToDevice));
	p->fpnDetermined = true;
	return OCTB200_OK;
}

int octb200_make_resample_curve(int n, float c0, float c1, float c2, float c3, float* out) {
	if (n < 4 || !out) return OCTB200_ERR_INVALID;
	curves::resample(n, c0, c1, c2, c3, out); return OCTB200_OK;
}
int octb200_make_dispersion_curve(int n, float d0, float d1, float d2, float d3, float* out) {
	if (n < 2 || !out) return OCTB200_ERR_INVALID;
	curves::dispersion(n, d0, d1, d2, d3, out); return OCTB200_OK;
}
int octb200_make_window_curve(int type, float center, float fill, int n, float* out) {
	if (n < 2 || !out || type < 0 || type > 5) return OCTB200_ERR_INVALID;
	curves::window(type, center, fill, n, out); return OCTB200_OK;
}
int octb200_make_sinusoidal_curve(int ascans, float* out) {
	if (ascans < 1 || !out) return OCTB200_ERR_INVALID;
	curves::sinusoidal(ascans, out); return OCTB200_OK;
}

/* ---------- host buffers ---------- */
int octb200_register_host_buffers(octb200_pipeline* p, void* h1, void* h2) {
	if (!p || !h1) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	if (p->hostRegistered) octb200_unregister_host_buffers(p);
	const size_t bytes = p->inBytes;
	CK(p, pin_host(h1, bytes, &p->hostBufMine[0]));
	p->hostBuf[0] = h1; p->hostBuf[1] = nullptr; p->hostBufMine[1] = false;
	if (h2 && h2 != h1) {
		cudaError_t e = pin_host(h2, bytes, &p->hostBufMine[1]);
		if (e != cudaSuccess) { if (p->hostBufMine[0]) cudaHostUnregister(h1); p->hostBuf[0] = nullptr; return fail(p, OCTB200_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
		p->hostBuf[1] = h2;
	}
	p->hostRegistered = true;
	return OCTB200_OK;
}
int octb200_unregister_host_buffers(octb200_pipeline* p) {
	if (!p) return OCTB200_ERR_INVALID;
	if (p->hostRegistered) {
		cudaStreamSynchronize(p->sH2D);
		for (int i = 0; i < 2; ++i) { if (p->hostBuf[i] && p->hostBufMine[i]) cudaHostUnregister(p->hostBuf[i]); p->hostBuf[i] = nullptr; p->hostBufMine[i] = false; }
		p->hostRegistered = false;
	}
	return OCTB200_OK;
}
static int reg_pair(octb200_pipeline* p, void* h1, void* h2, size_t bytes, void** dst, bool* mine, size_t* dstBytes, bool* flag) {
	if (!p || !h1 || !h2) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, pin_host(h1, bytes, &mine[0]));
	cudaError_t e = pin_host(h2, bytes, &mine[1]);
	if (e != cudaSuccess) { if (mine[0]) cudaHostUnregister(h1); return fail(p, OCTB200_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
	dst[0] = h1; dst[1] = h2; *dstBytes = bytes; *flag = true;
	return OCTB200_OK;
}
static void unreg_pair(octb200_pipeline* p, void** dst, bool* mine, bool* flag) {
	if (!*flag) return;
	cudaStreamSynchronize(p->sD2H);
	for (int i = 0; i < 2; ++i) { if (dst[i] && mine[i]) cudaHostUnregister(dst[i]); dst[i] = nullptr; mine[i] = false; }
	*flag = false;
}
int octb200_register_streaming_buffers(octb200_pipeline* p, void* h1, void* h2, size_t bytes) {
	if (!p) return OCTB200_ERR_INVALID;
	if (bytes < (size_t)(p->S / 2) * p->rawBytes) return fail(p, OCTB200_ERR_INVALID, "streaming buffer too small");
	unreg_pair(p, p->hostStream, p->hostStreamMine, &p->hostStreamRegistered);
	return reg_pair(p, h1, h2, bytes, p->hostStream, p->hostStreamMine, &p->hostStreamBytes, &p->hostStreamRegistered);
}
int octb200_unregister_streaming_buffers(octb200_pipeline* p) {
	if (!p) return OCTB200_ERR_INVALID;
	unreg_pair(p, p->hostStream, p->hostStreamMine, &p->hostStreamRegistered);
	return OCTB200_OK;
}
int octb200_register_float_streaming_buffers(octb200_pipeline* p, void* h1, void* h2, size_t bytes) {
	if (!p) return OCTB200_ERR_INVALID;
	if (bytes < (size_t)(p->S / 2) * sizeof(float)) return fail(p, OCTB200_ERR_INVALID, "float streaming buffer too small");
	unreg_pair(p, p->hostFloat, p->hostFloatMine, &p->hostFloatRegistered);
	return reg_pair(p, h1, h2, bytes, p->hostFloat, p->hostFloatMine, &p->hostFloatBytes, &p->hostFloatRegistered);
}
int octb200_unregister_float_streaming_buffers(octb200_pipeline* p) {
	if (!p) return OCTB200_ERR_INVALID;
	unreg_pair(p, p->hostFloat, p->hostFloatMine, &p->hostFloatRegistered);
	return OCTB200_OK;
}
int octb200_set_callbacks(octb200_pipeline* p, octb200_host_callback s, octb200_host_callback f, octb200_host_callback b) {
	if (!p) return OCTB200_ERR_INVALID;
	p->cbStreaming = s; p->cbFloat = f; p->cbBackground = b;
	return OCTB200_OK;
}

/* ---------- hot path ---------- */
/* after every chain that read internal raw slot `s` -- also a re-run on the resident buffer and a chain that failed half way: the next
   upload into that slot waits for the kernels enqueued so far */
static int chain_on_slot(octb200_pipeline* p, int s) {
	const int rc = run_chain(p, p->dRaw[s]);
	const cudaError_t e = cudaEventRecord(p->evRawFree[s], p->sCompute);
	if (rc) return rc;
	CK(p, e);
	return OCTB200_OK;
}
int octb200_process_host(octb200_pipeline* p, const void* hRaw) {
	if (!p) return OCTB200_ERR_INVALID;
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	if (!hRaw) {
		if (p->slot < 0 && !p->lastDeviceRaw) return fail(p, OCTB200_ERR_NOT_READY, "no buffer has been uploaded yet");
		return p->slot >= 0 ? chain_on_slot(p, p->slot) : run_chain(p, p->lastDeviceRaw);
	}
	const int s = (p->slot + 1) % (int)p->dRaw.size();
	const size_t bytes = p->inBytes;
	CK(p, cudaStreamWaitEvent(p->sH2D, p->evRawFree[s], 0));       /* kernels that still read this slot */
	CK(p, cudaMemcpyAsync(p->dRaw[s], hRaw, bytes, cudaMemcpyHostToDevice, p->sH2D));   /* cuda_code.cu:1404 */
	CK(p, cudaEventRecord(p->evRawReady[s], p->sH2D));
	CK(p, cudaStreamWaitEvent(p->sCompute, p->evRawReady[s], 0));
	p->slot = s;
	int rc = chain_on_slot(p, s);
	if (rc) return rc;
	/* the producer may overwrite hRaw as soon as we return (processing.cpp:191) -- same contract as cuda_code.cu:1416-1419 */
	CK(p, cudaEventSynchronize(p->evRawReady[s]));
	return OCTB200_OK;
}

int octb200_process_device(octb200_pipeline* p, const void* dRaw) {
	if (!p) return OCTB200_ERR_INVALID;
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	if (!dRaw) {
		if (!p->lastDeviceRaw && p->slot >= 0) return chain_on_slot(p, p->slot);     /* re-run on the internal slot the last upload filled */
		dRaw = p->lastDeviceRaw;
		if (!dRaw) return fail(p, OCTB200_ERR_NOT_READY, "no device buffer to re-process");
	}
	if (reinterpret_cast<uintptr_t>(dRaw) & 15) return fail(p, OCTB200_ERR_INVALID, "device raw pointer must be 16-byte aligned");
	p->lastDeviceRaw = dRaw;
	return run_chain(p, dRaw);
}

int octb200_sync(octb200_pipeline* p) {
	if (!p) return OCTB200_ERR_INVALID;
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaStreamSynchronize(p->sH2D));
	CK(p, cudaStreamSynchronize(p->sCompute));
	CK(p, cudaStreamSynchronize(p->sD2H));
	return OCTB200_OK;
}
uint32_t octb200_current_buffer_nr(const octb200_pipeline* p) { return p ? p->bufferNumberInVolume : 0; }

/* ---------- results ---------- */
float* octb200_output_device_ptr(octb200_pipeline* p, uint32_t nr) {
	if (!p || nr >= (uint32_t)p->V) return nullptr;
	return p->dVolume + (size_t)(p->S / 2) * nr;
}
int octb200_copy_output(octb200_pipeline* p, float* host, uint32_t nr) {
	if (!p || !host || nr >= (uint32_t)p->V) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaStreamSynchronize(p->sCompute));
	CK(p, cudaMemcpy(host, p->dVolume + (size_t)(p->S / 2) * nr, (size_t)(p->S / 2) * sizeof(float), cudaMemcpyDeviceToHost));
	return OCTB200_OK;
}
int octb200_bind_output(octb200_pipeline* p, void* dVolume) {
	if (!p) return OCTB200_ERR_INVALID;
	if (dVolume && (reinterpret_cast<uintptr_t>(dVolume) & 15)) return fail(p, OCTB200_ERR_INVALID, "volume pointer must be 16-byte aligned");
	p->dVolume = dVolume ? static_cast<float*>(dVolume) : p->dVolumeOwned;
	return OCTB200_OK;
}

int octb200_bscan_frame(octb200_pipeline* p, uint32_t frameNr, uint32_t nFrames, int fn, float* dOut) {
	if (!p || !dOut) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	const unsigned Btot = (unsigned)(p->B * p->V);
	if (frameNr >= Btot) frameNr = 0;                                /* cuda_code.cu:1278 */
	CK(p, launch_bscan_frame(dOut, p->dVolume, Btot, (unsigned)(p->H * p->A), frameNr, nFrames, fn, p->sCompute)); p->launches++;
	return OCTB200_OK;
}
int octb200_enface_frame(octb200_pipeline* p, uint32_t frameNr, uint32_t nFrames, int fn, float* dOut) {
	if (!p || !dOut) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	if (frameNr >= (unsigned)p->H) frameNr = 0;                      /* cuda_code.cu:1302 */
	CK(p, launch_enface_frame(dOut, p->dVolume, (unsigned)p->H, (unsigned)(p->A * p->B * p->V), frameNr, nFrames, fn, p->sCompute)); p->launches++;
	return OCTB200_OK;
}
int octb200_volume_u8(octb200_pipeline* p, uint32_t nr, uint8_t* dOut) {
	if (!p || !dOut || nr >= (uint32_t)p->V) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, launch_volume_u8(dOut, p->dVolume + (size_t)(p->S / 2) * nr, p->S / 2, nr, (unsigned)p->B, (unsigned)p->A, (unsigned)(p->B * p->V),
	                       (unsigned)p->H, p->smCount, p->sCompute)); p->launches++;
	return OCTB200_OK;
}
int octb200_float_to_output(octb200_pipeline* p, uint32_t nr, void* dOut) {
	if (!p || !dOut || nr >= (uint32_t)p->V) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, launch_float_to_output(dOut, p->dVolume + (size_t)(p->S / 2) * nr, (int)p->cfg.bitDepth, p->S / 2, p->smCount, p->sCompute)); p->launches++;
	return OCTB200_OK;
}

/* ---------- en-face gather over peer memory ---------- */
int octb200_enface_gather_init(octb200_pipeline* p, int rank, int world, uint32_t globalLines, uint32_t lineOffset, void* handleOut) {
	if (!p || !handleOut || world < 1 || world > OCT_MAX_PEERS || rank < 0 || rank >= world) return fail(p, OCTB200_ERR_INVALID, "bad rank/world (at most %d ranks)", OCT_MAX_PEERS);
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	const unsigned E = (unsigned)(p->A * p->B * p->V);
	if ((unsigned long long)lineOffset + E > globalLines) return fail(p, OCTB200_ERR_INVALID, "shard [%u, %u) exceeds the %u lines of the volume", lineOffset, lineOffset + E, globalLines);
	octb200_enface_gather_close(p);
	auto& g = p->eg;
	g.world = world; g.rank = rank; g.Eglobal = globalLines; g.offset = lineOffset; g.seq = 0; g.consumedSeq = 0;
	g.frameStride = ((size_t)globalLines * sizeof(float) + 255) / 256 * 256;
	{ int rc = dalloc(p, &g.window, OCT_GATHER_HEADER_BYTES + OCT_GATHER_FRAMES * g.frameStride); if (rc) return rc; }
	{ int rc = dalloc(p, &g.counter, 4); if (rc) return rc; }
	{ int rc = dalloc(p, &g.display, (size_t)globalLines + 4); if (rc) return rc; }
	cudaIpcMemHandle_t h;
	CK(p, cudaIpcGetMemHandle(&h, g.window));
	static_assert(sizeof(h) == OCTB200_IPC_HANDLE_BYTES, "ipc handle size");
	std::memcpy(handleOut, &h, sizeof(h));
	g.peerBase[rank] = g.window;
	return OCTB200_OK;
}
int octb200_enface_gather_connect(octb200_pipeline* p, const void* handles) {
	if (!p || !handles || !p->eg.window) return fail(p, OCTB200_ERR_INVALID, "enface_gather_init first");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	auto& g = p->eg;
	for (int r = 0; r < g.world; ++r) {
		if (r == g.rank || g.opened[r]) continue;
		cudaIpcMemHandle_t h;
		std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * OCTB200_IPC_HANDLE_BYTES, sizeof(h));
		void* base = nullptr;
		CK(p, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
		g.peerBase[r] = static_cast<unsigned char*>(base); g.opened[r] = true;
	}
	g.connected = true;
	return OCTB200_OK;
}
int octb200_enface_gather(octb200_pipeline* p, uint32_t frameNr, uint32_t nFrames, int fn) {
	if (!p || !p->eg.connected) return fail(p, OCTB200_ERR_NOT_READY, "en-face gather is not connected");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	const GatherDev d = next_gather(p, frameNr, nFrames, fn);
	CK(p, launch_gather_standalone(p, d)); p->launches++;
	CK(p, consume_gather(p));
	return OCTB200_OK;
}
int octb200_enface_gather_auto(octb200_pipeline* p, int enable, uint32_t frameNr, uint32_t nFrames, int fn) {
	if (!p) return OCTB200_ERR_INVALID;
	if (enable && !p->eg.connected) return fail(p, OCTB200_ERR_NOT_READY, "en-face gather is not connected");
	p->eg.autoOn = enable != 0; p->eg.autoFrame = frameNr; p->eg.autoFrames = nFrames; p->eg.autoFn = fn;
	return OCTB200_OK;
}
int octb200_enface_gather_wait(octb200_pipeline* p, float** dFrame) {
	if (!p || !p->eg.connected || p->eg.seq == 0) return fail(p, OCTB200_ERR_NOT_READY, "no en-face gather issued");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, consume_gather(p));          /* already enqueued behind the gather itself; a no-op then */
	if (dFrame) *dFrame = p->eg.display;
	return OCTB200_OK;
}
int octb200_enface_gather_status(octb200_pipeline* p, uint32_t* sequence, uint32_t* ackTimeouts, uint32_t* arrivalTimeouts) {
	if (!p || !p->eg.counter) return fail(p, OCTB200_ERR_NOT_READY, "enface_gather_init first");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	unsigned st[2] = { 0, 0 };
	CK(p, cudaStreamSynchronize(p->sCompute));
	CK(p, cudaMemcpy(st, p->eg.counter + 2, sizeof(st), cudaMemcpyDeviceToHost));
	if (sequence) *sequence = p->eg.seq;
	if (ackTimeouts) *ackTimeouts = st[0];
	if (arrivalTimeouts) *arrivalTimeouts = st[1];
	return OCTB200_OK;
}
int octb200_enface_gather_close(octb200_pipeline* p) {
	if (!p) return OCTB200_ERR_INVALID;
	auto& g = p->eg;
	if (!g.window && !g.counter) return OCTB200_OK;
	cudaSetDevice(p->device);
	if (p->sCompute) cudaStreamSynchronize(p->sCompute);
	for (int r = 0; r < OCT_MAX_PEERS; ++r) {
		if (g.opened[r] && g.peerBase[r]) cudaIpcCloseMemHandle(g.peerBase[r]);
		g.opened[r] = false; g.peerBase[r] = nullptr;
	}
	dfree(g.window); dfree(g.counter); dfree(g.display);
	g.connected = false; g.world = 0; g.seq = 0; g.consumedSeq = 0; g.autoOn = false;
	return OCTB200_OK;
}

/* ---------------- dispersion-estimator sweep ---------------- */
int octb200_dispersion_sweep(octb200_pipeline* p, const void* raw, const octb200_sweep_config* c, const float* coeffs, float* metricsOut, float* ascansOut) {
	if (!p || !raw || !c || !coeffs || !metricsOut) return fail(p, OCTB200_ERR_INVALID, "null argument");
	if (c->lines < 1 || c->trials < 1 || c->trials > 65535 || c->metric < 0 || c->metric > 3) return fail(p, OCTB200_ERR_INVALID, "bad sweep configuration");
	if (!p->regKernel || p->rawBytes != 2)
		return fail(p, OCTB200_ERR_INVALID, "the dispersion sweep runs on the fused kernel: 1024 or 2048 samples per line in a 16-bit container");
	if (c->logScale && !(c->logMax != c->logMin)) return fail(p, OCTB200_ERR_INVALID, "log scaling needs max != min");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	const octb200_params& q = p->prm;
	const Stage st = select_stage(p);
	if (q.resampling && !p->haveResample) return fail(p, OCTB200_ERR_NOT_READY, "resampling enabled but no resample curve set");
	if (q.windowing && !p->haveWindow) return fail(p, OCTB200_ERR_NOT_READY, "windowing enabled but no window curve set");
	const int N = p->N, H = p->H, K = (int)c->trials, L = (int)c->lines;

	/* launch shape first: a stage that does not fit the fused kernel cannot be swept */
	{
		int g = 0, t = 0, sm = 0;
		fused_launch_shape(p->R, st.sa, st.roll, SRC_RAW16, st.HB, st.HA, p->smCount, L, &g, &t, &sm);
		if (t < 32 * p->R) return fail(p, OCTB200_ERR_INVALID, "this rolling window / interpolation does not fit the fused kernel");
	}
	int rc;
	const size_t rawBytes = (size_t)L * N * 2;
	if ((rc = grow(p, p->dSweepRaw, p->sweepRawBytes, rawBytes + 64))) return rc;
	if ((rc = grow(p, p->dSweepLut, p->sweepLutElems, (size_t)K * 2 * N))) return rc;
	if ((rc = grow(p, p->dSweepOut, p->sweepOutElems, (size_t)K * L * H))) return rc;
	if (p->sweepPhaseElems < (size_t)K * N) { dfree(p->dSweepPhase); dfree(p->dSweepPhasor); p->sweepPhaseElems = 0; }
	if (!p->dSweepPhase) {
		if ((rc = dalloc(p, &p->dSweepPhase, (size_t)K * N))) return rc;
		if ((rc = dalloc(p, &p->dSweepPhasor, (size_t)K * N))) return rc;
		p->sweepPhaseElems = (size_t)K * N;
	}
	if ((rc = grow(p, p->dSweepMetric, p->sweepMetricElems, (size_t)K))) return rc;

	/* raw center A-scans: host or device memory */
	CK(p, cudaMemcpyAsync(p->dSweepRaw, raw, rawBytes, cudaMemcpyDefault, p->sCompute));

	/* trial phase curves = OctAlgorithmParameters::updateDispersionCurve with (d0, d1, d2_k, d3_k); phasors on the device like
	   fillDispersivePhase, then one stage-LUT image per trial */
	std::vector<float> phase((size_t)K * N);
	for (int k = 0; k < K; ++k) curves::dispersion(N, coeffs[4 * k], coeffs[4 * k + 1], coeffs[4 * k + 2], coeffs[4 * k + 3], phase.data() + (size_t)k * N);
	CK(p, cudaMemcpyAsync(p->dSweepPhase, phase.data(), sizeof(float) * K * N, cudaMemcpyHostToDevice, p->sCompute));
	launch_fill_phase(p->dSweepPhasor, p->dSweepPhase, K * N, p->sCompute); p->launches++;
	std::vector<float2> phasor((size_t)K * N);
	CK(p, cudaMemcpyAsync(phasor.data(), p->dSweepPhasor, sizeof(float2) * K * N, cudaMemcpyDeviceToHost, p->sCompute));
	CK(p, cudaStreamSynchronize(p->sCompute));
	const float* res = q.resampling ? p->hResample.data() : nullptr;
	const float* win = q.windowing ? p->hWindow.data() : nullptr;
	std::vector<float4> luts((size_t)K * 2 * N), one;
	for (int k = 0; k < K; ++k) {
		build_stage_luts_paired(N, p->R, q.resamplingInterpolation == OCTB200_INTERP_CUBIC ? 1 : 0, res, win, phasor.data() + (size_t)k * N, one,
		                        lut_taps_mode(p->R, st));
		std::memcpy(luts.data() + (size_t)k * 2 * N, one.data(), sizeof(float4) * 2 * N);
	}
	CK(p, cudaMemcpyAsync(p->dSweepLut, luts.data(), sizeof(float4) * luts.size(), cudaMemcpyHostToDevice, p->sCompute));

	/* ONE launch for all trials: gridDim.y = trials, every trial reads the same raw lines through its own LUT */
	FusedArgs fa{};
	fa.raw = reinterpret_cast<const uint16_t*>(p->dSweepRaw);
	fa.lutB = p->dSweepLut; fa.tw = p->dTw; fa.ctw = p->dCtw; fa.meanLine = p->dMeanLine; fa.ppbg = p->dPpbg;
	fa.totalSamples = (long long)L * N; fa.lines = L; fa.A = L;
	fa.flip = 0; fa.bscanBase = 0; fa.shiftBits = q.bitshift ? 4 : 0; fa.W = st.W; fa.HB = st.HB; fa.HA = st.HA;
	fa.out = p->dSweepOut;
	fa.trials = K; fa.trialLutStride = 2 * N; fa.trialOutStride = (long long)L * H;
	/* output in the units of the reference's CPU path (processor.tpp:424-470): IFFT normalised by 1/N;
	   log: coeff * ((10 log10(|x|^2 / N) - min) / (max - min) + addend); linear: |x| */
	EpiConsts e; std::memset(&e, 0, sizeof(e));
	e.logMode = c->logScale ? 1 : 0;
	if (c->logScale) {
		const double range = (double)c->logMax - (double)c->logMin;
		e.scaleA = (float)((double)c->logCoeff * 10.0 * std::log10(2.0) / range);
		e.scaleB = (float)((double)c->logCoeff * ((-30.0 * std::log10((double)N) - (double)c->logMin) / range + (double)c->logAddend));
	} else {
		e.scaleA = (float)(1.0 / (double)N); e.scaleB = 0.0f;
	}
	fa.epi = e;
	CK(p, launch_fused(p->R, st.sa, st.roll, SRC_RAW16, fa, p->smCount, p->sCompute)); p->launches++;
	CK(p, launch_sweep_metric(p->dSweepMetric, p->dSweepOut, K, L, H, c->metric, c->metricThreshold, c->samplesToIgnore, p->sCompute)); p->launches++;
	CK(p, cudaMemcpyAsync(metricsOut, p->dSweepMetric, sizeof(float) * K, cudaMemcpyDeviceToHost, p->sCompute));
	if (ascansOut) CK(p, cudaMemcpyAsync(ascansOut, p->dSweepOut, sizeof(float) * (size_t)K * L * H, cudaMemcpyDeviceToHost, p->sCompute));
	CK(p, cudaStreamSynchronize(p->sCompute));
	return OCTB200_OK;
}

/* ---------- timing ---------- */
void* octb200_compute_stream(octb200_pipeline* p) { return p ? (void*)p->sCompute : nullptr; }
int octb200_event_record(octb200_pipeline* p, int slot) {
	if (!p || slot < 0 || slot >= 8) return OCTB200_ERR_INVALID;
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaEventRecord(p->evTiming[slot], p->sCompute));
	return OCTB200_OK;
}
int octb200_event_elapsed_ms(octb200_pipeline* p, int a, int b, float* ms) {
	if (!p || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return OCTB200_ERR_INVALID;
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	CK(p, cudaEventSynchronize(p->evTiming[b]));
	CK(p, cudaEventElapsedTime(ms, p->evTiming[a], p->evTiming[b]));
	return OCTB200_OK;
}
uint64_t octb200_launch_count(const octb200_pipeline* p) { return p ? p->launches : 0; }

int octb200_time_kernel(octb200_pipeline* p, const void* dRaw, int iters, float* msPerIter) {
	if (!p || !dRaw || iters < 1 || !msPerIter) return fail(p, OCTB200_ERR_INVALID, "bad argument");
	if (use_device(p)) return fail(p, OCTB200_ERR_CUDA, "cudaSetDevice failed");
	const Stage st = select_stage(p);
	if (p->lutsDirty || st.sa != p->lutSa || st.roll != p->lutRoll) { int rc = rebuild_luts(p); if (rc) return rc; }
	const bool fpn = p->prm.fixedPatternNoiseRemoval != 0;
	float* slab = p->dVolume + (size_t)(p->S / 2) * p->bufferNumberInVolume;
	CK(p, cudaEventRecord(p->evTiming[6], p->sCompute));
	for (int i = 0; i < iters; ++i) {
		if (p->mode == OCTB200_FFT_CUFFT) {
			if (p->packed12) return fail(p, OCTB200_ERR_INVALID, "time_kernel: packed input is only timed on the direct fused path");
			PreArgs pa = pre_args(p, st, dRaw, p->lines);
			int rc = ensure_fft_buffer(p); if (rc) return rc;
			CK(p, launch_pre(pa, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute));
		} else if (p->mode == OCTB200_FFT_FUSED && !p->regKernel) {
			GenericArgs ga = generic_args(p, st, dRaw, p->lines);
			ga.out = slab; ga.epi = epi_for(p, fpn && p->fpnDetermined, false);
			CK(p, launch_generic(ga, p->rawBytes, st.sa, st.roll, p->smCount, p->sCompute));
		} else {
			const int container = p->rawBytes == 1 ? SRC_RAW8 : (p->rawBytes == 4 ? SRC_RAW32 : SRC_RAW16);
			if (p->mode == OCTB200_FFT_FUSED && container != SRC_RAW16 && (st.roll || st.sa == SA_LANCZOS))
				return fail(p, OCTB200_ERR_INVALID, "time_kernel: this stage of a u8 / u32 container runs on the split chain, not on the fused kernel");
			const int src = (p->mode == OCTB200_FFT_FUSED) ? ((p->packed12 && st.sa != SA_LANCZOS && !st.roll) ? SRC_RAW12P : container) : SRC_CPLX;
			if (p->packed12 && src != SRC_RAW12P) return fail(p, OCTB200_ERR_INVALID, "time_kernel: packed input is only timed on the direct fused path");
			if (src == SRC_CPLX) { int rc = ensure_fft_buffer(p); if (rc) return rc; }
			FusedArgs fa = fused_args(p, st, dRaw, p->lines);
			fa.out = slab; fa.epi = epi_for(p, fpn && p->fpnDetermined, false);
			fa.pdl = (i > 0 && src != SRC_CPLX && !(p->cfg.flags & OCTB200_FLAG_NO_DEPENDENT_LAUNCH)) ? 1 : 0;
			CK(p, launch_fused(p->R, st.sa, st.roll, src, fa, p->smCount, p->sCompute));
		}
		p->launches++;
	}
	CK(p, cudaEventRecord(p->evTiming[7], p->sCompute));
	CK(p, cudaEventSynchronize(p->evTiming[7]));
	float ms = 0.f;
	CK(p, cudaEventElapsedTime(&ms, p->evTiming[6], p->evTiming[7]));
	*msPerIter = ms / (float)iters;
	return OCTB200_OK;
}

}  // extern "C"
