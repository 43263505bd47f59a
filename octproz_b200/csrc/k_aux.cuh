/*
 * k_aux.cuh -- argument blocks and launchers of the kernels in k_aux.cu and k_fused_*.cu.
 */
#pragma once
#include "k_fused.cuh"

namespace octb200 {

/* generic pre-FFT kernel (any N, u8/u16/u32 container): raw -> float2 FFT input in HBM */
struct PreArgs {
	const void* raw;
	float2* out;            /* [lines][N] */
	const float4* lutB;     /* N entries, natural order (R = 1 layout), global memory */
	long long totalSamples;
	int lines;
	int N;
	int shiftBits;
	int W;
	int HB, HA;
	int useBulk;            /* 1: cp.async.bulk staging (16-byte aligned geometry), 0: plain loads */
};

/* generic fused kernel (k_generic.cu): any even N <= 8192 with prime factors in {2, 3, 5, 7, 11, 13}, u8 / u16 / u32 containers; the
 * transform runs as Stockham passes between two shared-memory line buffers */
struct GenericArgs {
	const void* raw;
	float* out;              /* [lines][N/2] processed output slab (flip folded into the line address) */
	float2* cplxOut;         /* != NULL: write the pre-FPN complex bins [lines][N/2] instead (FPN determination pass) */
	const float4* lutB;      /* N entries, natural order: { byte offset of tap n1, w cos, w sin, t } (Lanczos / no resampling) */
	const float4* lutG;      /* 4-tap interpolators: 2 N entries { byte offset of tap n1 - 1, w0, w1, w2 } { w3, w cos, w sin, - } */
	const float2* tw;        /* per-pass twiddle tables (generic_twiddle_layout) */
	const float2* meanLine;  /* N/2 */
	const float* ppbg;       /* N/2 */
	EpiConsts epi;
	long long totalSamples;
	int lines, N, A;
	int flip;
	unsigned bscanBase, flipEnd;
	int shiftBits, W, HB, HA;
	int useBulk;             /* 1: cp.async.bulk staging (16-byte aligned geometry), 0: plain loads */
	int nPass;
	int radix[16];
	int twOff[16];
	unsigned magic[16];      /* ceil(2^32 / Ns) of every pass: j mod Ns without a division */
	int twEntries;           /* total entries of the twiddle tables (copied into shared memory once per CTA) */
};
/* plan / twiddle builders: generic_fft.cuh (inline, shared with the CPU emulator) */
bool generic_fits(int N, int rawBytes, int HB, int HA, bool roll, int twEntries);
cudaError_t launch_generic(const GenericArgs& a, int rawBytes, int sa, bool roll, int smCount, cudaStream_t st);

/* post kernel after cuFFT: complex [lines][N] -> float [lines][N/2] */
struct PostArgs {
	const float2* in;
	float* out;
	const float2* meanLine;
	const float* ppbg;
	EpiConsts epi;
	int lines, N, A;
	int flip;
	unsigned bscanBase;
	unsigned flipEnd;       /* see FusedArgs::flipEnd */
};

/* en-face extraction + all-gather over peer memory */
struct EnfaceGatherArgs {
	float* frames[OCT_MAX_PEERS];      /* this sequence number's frame window of every rank (peer-mapped device pointers) */
	unsigned* flags[OCT_MAX_PEERS];    /* flag words of every rank; word [rank] is written by this rank */
	const float* vol;
	unsigned* counter;                 /* local CTA completion counter (zero between launches) */
	unsigned* status;                  /* local time-out counters (GatherDev::status) */
	unsigned W, E;                     /* depth bins per line, local lines */
	unsigned frameNr, nFrames; int fn;
	unsigned Eglobal, offset;          /* lines of the whole (sharded) volume, first line of this shard */
	int world, rank;
	unsigned seq;
};
/* consumer side of the gather: wait for all slabs of frame `seq`, copy the frame out, acknowledge to every producer */
struct EnfaceConsumeArgs {
	unsigned* peerHeaders[OCT_MAX_PEERS];   /* window header of every rank (ack[] at word OCT_GATHER_ACK) */
	const unsigned* window;                 /* this rank's window header (arrived[] at word 0) */
	const float* frame;                     /* frame buffer of this sequence number in this rank's window */
	float* display;                         /* private copy handed to the caller */
	unsigned* counter;                      /* local CTA completion counter of the consume kernel */
	unsigned* status;
	unsigned Eglobal, seq;
	int world, rank;
};
cudaError_t launch_enface_gather(const EnfaceGatherArgs& a, cudaStream_t st);
cudaError_t launch_enface_consume(const EnfaceConsumeArgs& a, int smCount, bool dependent, cudaStream_t st);

cudaError_t launch_sweep_metric(float* metrics, const float* data, int trials, int lines, int H, int metric, float thr, int ignore, cudaStream_t st);
cudaError_t launch_unpack12(uint16_t* out, const void* in, long long octets, int smCount, cudaStream_t st);
void launch_fill_phase(float2* ph, const float* phase, int n, cudaStream_t st);
cudaError_t launch_pre(const PreArgs& a, int rawBytes, int sa, bool roll, int smCount, cudaStream_t st);
cudaError_t launch_post(const PostArgs& a, int smCount, cudaStream_t st);
/* segStats: NULL or [9][bins] of { mean.x, mean.y, variance, mean power } -- every candidate segment of every bin */
cudaError_t launch_fpn_minvar(float2* meanLine, const float2* in, int bins, int stride, int height, float4* segStats, cudaStream_t st);
cudaError_t launch_sinusoidal(float* out, const float* in, const float* curve, int H, int A, long long samples,
                              int ppbgOn, const float* ppbg, float w, float o, unsigned short* conv, float convScale,
                              int smCount, cudaStream_t st);
cudaError_t launch_ppbg_record(float* bg, const float* data, int H, int A, cudaStream_t st);
cudaError_t launch_ppbg_remove(float* data, const float* bg, float w, float o, int H, long long samples, int smCount, cudaStream_t st);
cudaError_t launch_bscan_frame(float* disp, const float* vol, unsigned Btot, unsigned F, unsigned frameNr, unsigned nFrames, int fn, cudaStream_t st);
cudaError_t launch_enface_frame(float* disp, const float* vol, unsigned W, unsigned E, unsigned frameNr, unsigned nFrames, int fn, cudaStream_t st);
cudaError_t launch_volume_u8(uint8_t* tex, const float* buf, long long samples, unsigned bufferNr, unsigned B, unsigned A, unsigned Btot,
                             unsigned depth, int smCount, cudaStream_t st);
cudaError_t launch_float_to_output(void* out, const float* in, int bitDepth, long long n, int smCount, cudaStream_t st);

/* fused / own-FFT kernel.  R = N/1024 (1 or 2).  Returns cudaErrorInvalidConfiguration if the shared
 * memory budget cannot hold even one line group (huge rolling window). */
cudaError_t launch_fused(int R, int sa, bool roll, int src, const FusedArgs& a, int smCount, cudaStream_t st);
/* the CTA shape launch_fused will use (for reports): */
void fused_launch_shape(int R, int sa, bool roll, int src, int HB, int HA, int smCount, int lines, int* grid, int* threads, int* smem);

}  // namespace octb200
