"""Headless mirror of the reference's acquisition side and processing loop (no Qt).

  AcquisitionParams / AcquisitionBuffer   octproz_devkit/src/acquisitionparameter.h:31-37, acquisitionbuffer.{h,cpp}
  AcquisitionSystem                        octproz_devkit/src/acquisitionsystem.h:58-73 (startAcquisition / stopAcquisition,
                                           public `buffer`, `params`, `acqusitionRunning` -- the reference's spelling)
  VirtualOCTSystem                         octproz_plugins/octproz-virtual-oct-system/src/virtualoctsystem.cpp:59-224
                                           (headerless little-endian raw file replay into a 2-slot buffer)
  Processing                               octproz/src/processing.cpp:136-229 (poll the double buffer, call the pipeline, release the
                                           buffer, volumes/buffers/B-scans/A-scans per second like the sidebar, :194-207)
  Recorder                                 octproz/src/recorder.cpp:99-152 (N buffers -> one headerless .raw file)

The handshake is the reference's: `bufferReadyArray[i]` set by the producer, cleared by the consumer, `currIndex` = last
filled slot.  Here producer and consumer run in two Python threads (the reference uses two QThreads).
"""
from __future__ import annotations

import math
import os
import threading
import time
from dataclasses import dataclass

import numpy as np

from .synth import container_dtype


@dataclass
class AcquisitionParams:
    samplesPerLine: int = 0
    ascansPerBscan: int = 0
    bscansPerBuffer: int = 0
    buffersPerVolume: int = 0
    bitDepth: int = 0


class AcquisitionBuffer:
    """two (or more) 128-byte aligned host buffers + ready flags (acquisitionbuffer.cpp:43-76)"""

    def __init__(self):
        self.bufferArray: list[np.ndarray] = []
        self.bufferReadyArray: list[bool] = []
        self.currIndex = -1
        self.bufferCnt = 0
        self.bytesPerBuffer = 0
        self._backing = []

    def allocateMemory(self, bufferCnt: int, bytesPerBuffer: int) -> bool:
        self.releaseMemory()
        self.bufferCnt, self.bytesPerBuffer = bufferCnt, bytesPerBuffer
        for _ in range(bufferCnt):
            raw = np.zeros(bytesPerBuffer + 128, np.uint8)
            off = (-raw.ctypes.data) % 128                       # posix_memalign(.., 128, ..)
            self._backing.append(raw)
            self.bufferArray.append(raw[off:off + bytesPerBuffer])
            self.bufferReadyArray.append(False)
        return True

    def releaseMemory(self) -> None:
        self.bufferArray, self.bufferReadyArray, self._backing = [], [], []
        self.currIndex = -1


class AcquisitionSystem:
    def __init__(self):
        self.buffer = AcquisitionBuffer()
        self.params = AcquisitionParams()
        self.acqusitionRunning = False          # sic (acquisitionsystem.h:66)
        self.on_acquisition_started = None      # signal acquisitionStarted(AcquisitionSystem*)
        self.on_acquisition_stopped = None

    def startAcquisition(self) -> None:
        raise NotImplementedError

    def stopAcquisition(self) -> None:
        self.acqusitionRunning = False


class VirtualOCTSystem(AcquisitionSystem):
    """file replay (virtualoctsystem.cpp).  Settings keys = virtualoctsystemsettingsdialog.h:27-38."""

    def __init__(self, file_path: str, bit_depth: int, width: int, height: int, depth: int, buffers_per_volume: int = 1,
                 buffers_from_file: int = 2, bscan_offset: int = 0, wait_time_us: int = 0, sync_with_processing: bool = True,
                 packed12: bool = False):
        """packed12: the file holds 12-bit samples packed two per three bytes (octproz_b200/packing.py) instead of containers --
        an extension over the reference's raw format; pair it with OctPipeline(input_packing=PACK_12P)"""
        super().__init__()
        self.packed12 = packed12
        self.file_path, self.bscan_offset, self.wait_time_us = file_path, bscan_offset, wait_time_us
        self.buffers_from_file, self.sync_with_processing = buffers_from_file, sync_with_processing
        self.params = AcquisitionParams(width, height, depth, buffers_per_volume, bit_depth)
        self.buffers_delivered = 0

    def init(self) -> bool:
        if not os.path.isfile(self.file_path):
            return False
        p = self.params
        return self.buffer.allocateMemory(2, self._bytes(p.samplesPerLine * p.ascansPerBscan * p.bscansPerBuffer))

    def _bytes(self, samples: int) -> int:
        if self.packed12:
            return samples * 3 // 2
        return samples * int(math.ceil(self.params.bitDepth / 8.0))

    def startAcquisition(self) -> None:
        if not self.init():
            if self.on_acquisition_stopped:
                self.on_acquisition_stopped()
            return
        p = self.params
        n_bytes = self._bytes(p.bscansPerBuffer * p.samplesPerLine * p.ascansPerBscan)
        offset = self._bytes(self.bscan_offset * p.samplesPerLine * p.ascansPerBscan)      # virtualoctsystem.cpp:167
        if self.buffers_from_file > 2:
            # acqcuisitionSimulationLargeFile / acquisitionSimulationWithMultiFileBuffers (virtualoctsystem.cpp:107-113, 226-290): successive
            # buffers of the file are streamed into the two acquisition buffers in turn, rewinding after `buffers_from_file` buffers
            self._stream_large_file(n_bytes, offset)
            if self.on_acquisition_stopped:
                self.on_acquisition_stopped()
            return
        with open(self.file_path, "rb") as f:
            f.seek(offset)
            b0 = f.read(n_bytes)
            f.seek(offset + (n_bytes if self.buffers_from_file == 2 else 0))                # :175-179
            b1 = f.read(n_bytes)
        for dst, src in ((self.buffer.bufferArray[0], b0), (self.buffer.bufferArray[1], b1)):
            dst[: len(src)] = np.frombuffer(src, np.uint8)
        self.acqusitionRunning = True
        self.buffer.currIndex = 1
        if self.on_acquisition_started:
            self.on_acquisition_started(self)
        buf = self.buffer
        while self.acqusitionRunning:                                                        # :196-223
            while self.sync_with_processing and buf.bufferReadyArray[buf.currIndex] and self.acqusitionRunning:
                time.sleep(0)
            nxt = (buf.currIndex + 1) % 2
            buf.currIndex = nxt
            if not buf.bufferReadyArray[nxt]:
                buf.bufferReadyArray[nxt] = True
                self.buffers_delivered += 1
            if self.wait_time_us > 0:
                time.sleep(self.wait_time_us * 1e-6)
        if self.on_acquisition_stopped:
            self.on_acquisition_stopped()


    def _stream_large_file(self, n_bytes: int, offset: int) -> None:
        buf = self.buffer
        with open(self.file_path, "rb") as f:
            f.seek(offset)
            read_buffers = 0
            self.acqusitionRunning = True
            buf.currIndex = 1
            nxt = 0
            if self.on_acquisition_started:
                self.on_acquisition_started(self)
            while self.acqusitionRunning:                                                    # virtualoctsystem.cpp:249-288
                while self.sync_with_processing and buf.bufferReadyArray[buf.currIndex] and self.acqusitionRunning:
                    time.sleep(0)
                if not buf.bufferReadyArray[nxt]:
                    data = f.read(n_bytes)
                    buf.bufferArray[nxt][: len(data)] = np.frombuffer(data, np.uint8)
                    read_buffers += 1
                    if read_buffers >= self.buffers_from_file:                               # rewind (:268-272)
                        f.seek(offset)
                        read_buffers = 0
                    buf.currIndex = nxt
                    buf.bufferReadyArray[nxt] = True
                    self.buffers_delivered += 1
                    nxt = (buf.currIndex + 1) % 2
                if self.wait_time_us > 0:
                    time.sleep(self.wait_time_us * 1e-6)


class Recorder:
    """Recorder::slot_record (recorder.cpp:99-152): append buffers to one headerless file, stop after `buffers_to_record`"""

    def __init__(self, path: str, buffers_to_record: int):
        self.path, self.buffers_to_record, self.recorded = path, buffers_to_record, 0
        self._f = open(path, "wb")

    def record(self, buf: np.ndarray) -> bool:
        if self.recorded >= self.buffers_to_record:
            return False
        self._f.write(np.ascontiguousarray(buf).tobytes())
        self.recorded += 1
        if self.recorded == self.buffers_to_record:
            self._f.close()
        return True


class Processing:
    """Processing::slot_start (processing.cpp:136-229) without Qt: `pipeline` is an OctPipeline (or a stand-in with the same
    initializeCuda / octCudaPipeline / sync / cleanupCuda methods)."""

    def __init__(self, pipeline, oct_params):
        self.pipeline, self.octParams = pipeline, oct_params
        self.stats = {}
        self.on_raw_data = None          # signal rawData(ptr, bitDepth, N, A, B, buffersPerVolume, currentBufferNr) (processing.h:110)
        self.processed_buffers = 0

    def slot_start(self, system: AcquisitionSystem, max_buffers: int | None = None) -> bool:
        """max_buffers: headless runs stop the acquisition after this many processed buffers (the GUI's Stop button)"""
        buf = system.buffer
        for i in range(len(buf.bufferReadyArray)):                   # blockBuffersForAcquisitionSystem (:124-128)
            buf.bufferReadyArray[i] = True
        q = self.octParams
        if not self.pipeline.initializeCuda(buf.bufferArray[0], buf.bufferArray[1], q):     # :151
            for i in range(len(buf.bufferReadyArray)):
                buf.bufferReadyArray[i] = False
            system.stopAcquisition()                                  # initializationFailed -> slot_stop (octprozapp.cpp:54)
            return False
        curr_nr = q.buffersPerVolume - 1
        for i in range(len(buf.bufferReadyArray)):                   # unblock (:130-134)
            buf.bufferReadyArray[i] = False
        t0 = time.perf_counter()
        n = 0
        while system.acqusitionRunning:                               # :176-218
            pos = buf.currIndex
            if pos >= 0 and buf.bufferReadyArray[pos]:
                curr_nr = (curr_nr + 1) % q.buffersPerVolume
                if self.on_raw_data:
                    self.on_raw_data(buf.bufferArray[pos], q.bitDepth, q.samplesPerLine, q.ascansPerBscan, q.bscansPerBuffer, q.buffersPerVolume, curr_nr)
                self.pipeline.octCudaPipeline(buf.bufferArray[pos])  # :187
                buf.bufferReadyArray[pos] = False                     # :191
                n += 1
                if max_buffers is not None and n >= max_buffers:
                    system.stopAcquisition()
            else:
                time.sleep(0)
        self.pipeline.sync()
        dt = time.perf_counter() - t0
        self.processed_buffers = n
        bps = n / dt if dt > 0 else 0.0
        self.stats = {"buffers_per_s": bps, "volumes_per_s": bps / q.buffersPerVolume, "bscans_per_s": bps * q.bscansPerBuffer,
                      "ascans_per_s": bps * q.bscansPerBuffer * q.ascansPerBscan,                                  # :198-201
                      "buffer_MB": buf.bytesPerBuffer / 1048576.0, "MB_per_s": bps * buf.bytesPerBuffer / 1048576.0}
        return True


def replay(file_path: str, oct_params, pipeline, buffers: int = 16, **vos_kwargs):
    """run `buffers` buffers of a raw file through the pipeline with the reference's thread structure; returns Processing.stats"""
    q = oct_params
    vos = VirtualOCTSystem(file_path, q.bitDepth, q.samplesPerLine, q.ascansPerBscan, q.bscansPerBuffer, q.buffersPerVolume, **vos_kwargs)
    proc = Processing(pipeline, q)
    started = threading.Event()
    vos.on_acquisition_started = lambda s: started.set()
    t = threading.Thread(target=vos.startAcquisition, daemon=True)
    t.start()
    if not started.wait(30):
        raise RuntimeError("virtual OCT system did not start (file missing?)")
    proc.slot_start(vos, max_buffers=buffers)
    vos.stopAcquisition()
    t.join(30)
    return proc


def write_raw_file(path: str, volume: np.ndarray) -> None:
    """headerless little-endian containers, the format the Virtual OCT System reads (docs/docs/faq.md:5)"""
    np.ascontiguousarray(volume).astype(volume.dtype.newbyteorder("<"), copy=False).tofile(path)


def read_raw_file(path: str, bit_depth: int, n: int, a: int, b: int, bscan_offset: int = 0) -> np.ndarray:
    dt = np.dtype(container_dtype(bit_depth)).newbyteorder("<")
    cnt = n * a * b
    return np.fromfile(path, dt, count=cnt, offset=bscan_offset * n * a * dt.itemsize).reshape(b, a, n)
