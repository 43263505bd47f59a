"""Host object that drives the C ABI the way the reference's `Processing` drives kernels.h.

`OctPipeline` keeps the reference's entry-point names for the path
(octproz/src/kernels.h:63-84, called from octproz/src/processing.cpp:151,187,227,276-311):
`initializeCuda`, `octCudaPipeline`, `cleanupCuda`, `cuda_register*StreamingBuffers`,
`changeDisplayed{Bscan,EnFace}Frame` -- so the parity tests read like the reference's call sites.
It owns no arithmetic: every call lands in liboctb200.so.  PyTorch is used only by callers for
device memory / NCCL; this module needs numpy + ctypes only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .params import OctAlgorithmParameters


def _ptr(x) -> int:
    """device or host address of a numpy array / torch tensor / int"""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if isinstance(x, np.ndarray):
        return int(x.ctypes.data)
    raise TypeError(f"cannot take the address of {type(x)}")


def device_copy(dst, src_ptr: int, nbytes: int) -> None:
    """copy `nbytes` from a raw device address (e.g. the frame octb200_enface_gather_wait returns) into a torch CUDA tensor; the raw
    memory is wrapped through the CUDA array interface, the copy runs on torch's current stream (synchronise the pipeline first)"""
    import torch

    class _Raw:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(src_ptr), False), "version": 2}
    src = torch.as_tensor(_Raw(), device=dst.device)
    dst.view(torch.uint8).reshape(-1)[: int(nbytes)].copy_(src)


class OctPipeline:
    def __init__(self, fft_mode: int = _lib.FFT_AUTO, device: int = -1, raw_slots: int = 2, bscan_index_base: int = 0,
                 input_packing: int = _lib.PACK_CONTAINER, flags: int = 0, bscans_in_unsharded_buffer: int = 0):
        """input_packing = PACK_12P: the raw buffers hold 12-bit samples packed two per three bytes (octproz_b200.packing.pack12),
        an extension over the reference's container formats.  flags: _lib.FLAG_* (octb200_config.flags).
        bscans_in_unsharded_buffer: shards only -- B-scans per buffer of the un-sharded acquisition (octb200_config, B-scan flip)"""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._fft_mode, self._device, self._raw_slots, self._bscan_base = fft_mode, device, raw_slots, bscan_index_base
        self._packing = input_packing
        self._flags = flags
        self._unsharded = bscans_in_unsharded_buffer
        self.params: OctAlgorithmParameters | None = None
        self._callbacks = None

    # ------------------------------------------------------------------ helpers
    def _ck(self, rc: int, what: str) -> None:
        if rc != _lib.OK:
            msg = self._lib.octb200_last_error(self._h if self._h else None)
            raise _lib.Octb200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def handle(self):
        return self._h

    @property
    def fft_mode(self) -> int:
        return self._lib.octb200_effective_fft_mode(self._h)

    # ------------------------------------------------------------------ kernels.h names
    def initializeCuda(self, h_buffer1, h_buffer2, params: OctAlgorithmParameters) -> bool:
        """kernels.h:63 / cuda_code.cu:1067-1162.  h_buffer1/2: the plugin's two host acquisition buffers
        (numpy arrays or None); they are pinned like the reference does (cuda_code.cu:1135-1136)."""
        cfg = _lib.Config(int(params.samplesPerLine), int(params.ascansPerBscan), int(params.bscansPerBuffer),
                          int(params.buffersPerVolume), int(params.bitDepth), int(self._device), int(self._raw_slots),
                          int(self._fft_mode), int(self._bscan_base), int(self._packing), int(self._flags), int(self._unsharded))
        rc = self._lib.octb200_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.OK:
            self._h = C.c_void_p()
            self._create_error = (self._lib.octb200_last_error(None) or b"").decode()
            return False
        self.params = params
        if h_buffer1 is not None:
            self._ck(self._lib.octb200_register_host_buffers(self._h, _ptr(h_buffer1), _ptr(h_buffer2)), "register_host_buffers")
        self.push_params(force_curves=True)
        return True

    def push_params(self, force_curves: bool = False) -> None:
        """marshal the parameter object into the POD and honour the *Updated edge triggers (cuda_code.cu:1433-1445)"""
        q = self.params
        cp = q.to_c()
        self._ck(self._lib.octb200_set_params(self._h, C.byref(cp)), "set_params")
        if q.resampling and (q.resamplingUpdated or force_curves):
            if q.resampleCurve is None:
                q.updateResampleCurve()
            self._ck(self._lib.octb200_set_resample_curve(self._h, q.resampleCurve.ctypes.data, len(q.resampleCurve)), "set_resample_curve")
            q.resamplingUpdated = False
        if q.dispersionCompensation and (q.dispersionUpdated or force_curves):
            if q.dispersionCurve is None:
                q.updateDispersionCurve()
            self._ck(self._lib.octb200_set_dispersion_curve(self._h, q.dispersionCurve.ctypes.data, len(q.dispersionCurve)), "set_dispersion_curve")
            q.dispersionUpdated = False
        if q.windowing and (q.windowUpdated or force_curves):
            if q.windowCurve is None:
                q.updateWindowCurve()
            self._ck(self._lib.octb200_set_window_curve(self._h, q.windowCurve.ctypes.data, len(q.windowCurve)), "set_window_curve")
            q.windowUpdated = False
        if q.postProcessBackgroundRemoval and q.postProcessBackgroundUpdated and q.postProcessBackground is not None:
            bg = np.ascontiguousarray(q.postProcessBackground, np.float32)
            self._ck(self._lib.octb200_set_postprocess_background(self._h, bg.ctypes.data, len(bg)), "set_postprocess_background")
            q.postProcessBackgroundUpdated = False
        # edge triggers are consumed by the pipeline (cuda_code.cu:1524,1561)
        q.redetermineFixedPatternNoise = False
        q.postProcessBackgroundRecordingRequested = False

    def octCudaPipeline(self, h_inputSignal) -> None:
        """kernels.h:64 / cuda_code.cu:1389-1605: one raw buffer from HOST memory (None re-processes the last one)"""
        self.push_params()
        self._ck(self._lib.octb200_process_host(self._h, _ptr(h_inputSignal)), "process_host")

    def process_device(self, d_raw) -> None:
        """device-resident variant: raw buffer already in HBM (torch tensor or address)"""
        self.push_params()
        self._ck(self._lib.octb200_process_device(self._h, _ptr(d_raw)), "process_device")

    def sync(self) -> None:
        self._ck(self._lib.octb200_sync(self._h), "sync")

    def cleanupCuda(self) -> None:
        if self._h:
            self._lib.octb200_destroy(self._h)
            self._h = C.c_void_p()

    def cuda_registerStreamingBuffers(self, h1, h2, nbytes: int) -> None:
        self._ck(self._lib.octb200_register_streaming_buffers(self._h, _ptr(h1), _ptr(h2), nbytes), "register_streaming_buffers")

    def cuda_unregisterStreamingBuffers(self) -> None:
        self._ck(self._lib.octb200_unregister_streaming_buffers(self._h), "unregister_streaming_buffers")

    def cuda_registerFloatStreamingBuffers(self, h1, h2, nbytes: int) -> None:
        self._ck(self._lib.octb200_register_float_streaming_buffers(self._h, _ptr(h1), _ptr(h2), nbytes), "register_float_streaming_buffers")

    def cuda_unregisterFloatStreamingBuffers(self) -> None:
        self._ck(self._lib.octb200_unregister_float_streaming_buffers(self._h), "unregister_float_streaming_buffers")

    def set_callbacks(self, streaming=None, float_streaming=None, background=None) -> None:
        """Gpu2HostNotifier::{dh2StreamingCallback, dh2FloatStreamingCallback, backgroundSignalCallback} (gpu2hostnotifier.h:47-49)"""
        wrap = lambda f: _lib.HOST_CALLBACK(f) if f else _lib.HOST_CALLBACK(0)
        self._callbacks = (wrap(streaming), wrap(float_streaming), wrap(background))   # keep alive
        self._ck(self._lib.octb200_set_callbacks(self._h, *self._callbacks), "set_callbacks")

    def changeDisplayedBscanFrame(self, frameNr: int, displayFunctionFrames: int, displayFunction: int, d_out) -> None:
        self._ck(self._lib.octb200_bscan_frame(self._h, frameNr, displayFunctionFrames, displayFunction, _ptr(d_out)), "bscan_frame")

    def changeDisplayedEnFaceFrame(self, frameNr: int, displayFunctionFrames: int, displayFunction: int, d_out) -> None:
        self._ck(self._lib.octb200_enface_frame(self._h, frameNr, displayFunctionFrames, displayFunction, _ptr(d_out)), "enface_frame")

    # ------------------------------------------------------------------ multi-GPU en-face gather over peer memory (include/octb200.h)
    def enface_gather_init(self, rank: int, world: int, global_lines: int, line_offset: int) -> bytes:
        """allocate this rank's frame window; returns its 64-byte IPC handle (exchange with torch.distributed, then connect)"""
        h = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        self._ck(self._lib.octb200_enface_gather_init(self._h, rank, world, global_lines, line_offset, h), "enface_gather_init")
        return bytes(h)

    def enface_gather_connect(self, handles: bytes) -> None:
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        self._ck(self._lib.octb200_enface_gather_connect(self._h, buf), "enface_gather_connect")

    def enface_gather(self, frameNr: int, displayFunctionFrames: int, displayFunction: int) -> None:
        """changeDisplayedEnFaceFrame of the whole sharded volume: extraction + P2P stores into every rank's frame, one kernel"""
        self._ck(self._lib.octb200_enface_gather(self._h, frameNr, displayFunctionFrames, displayFunction), "enface_gather")

    def enface_gather_auto(self, enable: bool, frameNr: int = 0, displayFunctionFrames: int = 1, displayFunction: int = 0) -> None:
        """every octCudaPipeline / process_device call also gathers this frame (inside the fused kernel's epilogue when possible)"""
        self._ck(self._lib.octb200_enface_gather_auto(self._h, int(enable), frameNr, displayFunctionFrames, displayFunction), "enface_gather_auto")

    def enface_gather_status(self) -> dict:
        """sequence number of the latest gather and the device-side time-out counters (0 / 0 in a healthy run); synchronises"""
        seq, a, b = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._ck(self._lib.octb200_enface_gather_status(self._h, C.byref(seq), C.byref(a), C.byref(b)), "enface_gather_status")
        return {"sequence": int(seq.value), "ack_timeouts": int(a.value), "arrival_timeouts": int(b.value)}

    def enface_gather_wait(self) -> int:
        """device address of this rank's display frame: the latest gathered frame, consumed (all ranks' slabs waited for, copied out of
        the window, acknowledged) by the kernel that every gather enqueues behind itself"""
        out = C.c_void_p()
        self._ck(self._lib.octb200_enface_gather_wait(self._h, C.byref(out)), "enface_gather_wait")
        return int(out.value or 0)

    def enface_gather_close(self) -> None:
        self._ck(self._lib.octb200_enface_gather_close(self._h), "enface_gather_close")

    # ------------------------------------------------------------------ dispersion-estimator sweep (include/octb200.h)
    def dispersion_sweep(self, raw, coeffs, metric: int, threshold: float, samples_to_ignore: int, log_scale: bool,
                         log_min: float = 0.0, log_max: float = 100.0, log_coeff: float = 1.0, log_addend: float = 0.0,
                         want_ascans: bool = False):
        """all trial coefficient sets {d0, d1, d2, d3} on the same A-scans in ONE launch of the fused kernel + one metric kernel
        (replaces the per-trial CPU re-run of dispersionestimationengine.cpp:118-158).  raw: [lines][N] u16 (numpy or device tensor).
        Returns metrics [trials] (and the A-scans [trials][lines][N/2] in the CPU path's units when asked)."""
        self.push_params()
        n = int(self.params.samplesPerLine)
        co = np.ascontiguousarray(coeffs, np.float32).reshape(-1, 4)
        if isinstance(raw, np.ndarray):
            raw = np.ascontiguousarray(raw)
            lines = raw.size // n
        else:
            lines = int(raw.numel()) // n
        cfg = _lib.SweepConfig(lines=lines, trials=co.shape[0], metric=int(metric), metricThreshold=float(threshold),
                               samplesToIgnore=int(samples_to_ignore), logScale=int(bool(log_scale)), logMin=float(log_min),
                               logMax=float(log_max), logCoeff=float(log_coeff), logAddend=float(log_addend))
        metrics = np.empty(co.shape[0], np.float32)
        ascans = np.empty((co.shape[0], lines, n // 2), np.float32) if want_ascans else None
        self._ck(self._lib.octb200_dispersion_sweep(self._h, _ptr(raw), C.byref(cfg), co.ctypes.data, metrics.ctypes.data,
                                                    ascans.ctypes.data if ascans is not None else None), "dispersion_sweep")
        return (metrics, ascans) if want_ascans else metrics

    # ------------------------------------------------------------------ results
    def output_ptr(self, buffer_nr: int = 0) -> int:
        return int(self._lib.octb200_output_device_ptr(self._h, buffer_nr) or 0)

    def bind_output(self, d_volume) -> None:
        self._ck(self._lib.octb200_bind_output(self._h, _ptr(d_volume)), "bind_output")

    def copy_output(self, buffer_nr: int = 0) -> np.ndarray:
        q = self.params
        out = np.empty((q.bscansPerBuffer, q.ascansPerBscan, q.samplesPerLine // 2), np.float32)
        self._ck(self._lib.octb200_copy_output(self._h, out.ctypes.data, buffer_nr), "copy_output")
        return out

    def current_buffer_nr(self) -> int:
        """slab of the volume the last process call wrote (bufferNumberInVolume, cuda_code.cu:1530-1535)"""
        return int(self._lib.octb200_current_buffer_nr(self._h))

    def volume_u8(self, buffer_nr: int, d_out) -> None:
        self._ck(self._lib.octb200_volume_u8(self._h, buffer_nr, _ptr(d_out)), "volume_u8")

    def float_to_output(self, buffer_nr: int, d_out) -> None:
        self._ck(self._lib.octb200_float_to_output(self._h, buffer_nr, _ptr(d_out)), "float_to_output")

    def fpn_mean_line(self) -> np.ndarray:
        n = int(self.params.samplesPerLine)
        out = np.empty((n, 2), np.float32)
        self._ck(self._lib.octb200_get_fpn_mean_line(self._h, out.ctypes.data, n), "get_fpn_mean_line")
        return out

    def fpn_segment_stats(self):
        """the nine candidate segments of every depth bin at the last fixed-pattern-noise determination:
        (stats [9][N/2][4] = mean.re, mean.im, fp32 variance, mean power; segment length L)"""
        h = int(self.params.samplesPerLine) // 2
        out = np.empty((9, h, 4), np.float32)
        seg = C.c_int()
        self._ck(self._lib.octb200_get_fpn_segment_stats(self._h, out.ctypes.data, h, C.byref(seg)), "get_fpn_segment_stats")
        return out, int(seg.value)

    def set_fpn_mean_line(self, re_im: np.ndarray) -> None:
        a = np.ascontiguousarray(re_im, np.float32)
        self._ck(self._lib.octb200_set_fpn_mean_line(self._h, a.ctypes.data, a.shape[0]), "set_fpn_mean_line")

    def postprocess_background(self) -> np.ndarray:
        n = int(self.params.samplesPerLine) // 2
        out = np.empty(n, np.float32)
        self._ck(self._lib.octb200_get_postprocess_background(self._h, out.ctypes.data, n), "get_postprocess_background")
        return out

    # ------------------------------------------------------------------ timing
    def event_record(self, slot: int) -> None:
        self._ck(self._lib.octb200_event_record(self._h, slot), "event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._ck(self._lib.octb200_event_elapsed_ms(self._h, a, b, C.byref(ms)), "event_elapsed_ms")
        return float(ms.value)

    def launch_count(self) -> int:
        return int(self._lib.octb200_launch_count(self._h))

    def time_kernel(self, d_raw, iters: int) -> float:
        ms = C.c_float()
        self._ck(self._lib.octb200_time_kernel(self._h, _ptr(d_raw), iters, C.byref(ms)), "time_kernel")
        return float(ms.value)

    def __del__(self):
        try:
            self.cleanupCuda()
        except Exception:
            pass
