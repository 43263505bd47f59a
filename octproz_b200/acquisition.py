"""Headless mirror of the reference's acquisition side and processing loop (no Qt).

  AcquisitionParams / AcquisitionBuffer   octproz_devkit/src/acquisitionparameter.h:31-37, acquisitionbuffer.{h,cpp}
  AcquisitionSystem                        octproz_devkit/src/acquisitionsystem.h:58-73 (startAcquisition / stopAcquisition,
                                           public `buffer`, `params`, `acqusitionRunning` -- the reference's spelling)
  VirtualOCTSystem                         octproz_plugins/octproz-virtual-oct-system/src/virtualoctsystem.cpp:59-224
                                           (headerless little-endian raw file replay into a 2-slot buffer)
  Processing                               octproz/src/processing.cpp:136-229 (poll the double buffer, call the pipeline, release the
                                           buffer, volumes/buffers/B-scans/A-scans per second like the sidebar, :194-207)
  Recorder / RecordingParams               octproz/src/recorder.cpp, octalgorithmparameters.h:84-98 (N buffers -> one headerless .raw file:
                                           file naming, start with the first buffer of a volume, abort, meta file = settings INI copy)

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


@dataclass
class RecordingParams:
    """OctAlgorithmParameters::RecordingParams (octalgorithmparameters.h:84-98) without the GUI-only screenshot switch"""
    timestamp: str = ""
    fileName: str = ""
    savePath: str = ""
    bufferSizeInBytes: int = 0
    buffersToRecord: int = 0
    startWithFirstBuffer: bool = False
    recordRaw: bool = False
    recordProcessed: bool = False
    saveMetaData: bool = False
    saveAs32bitFloat: bool = False
    stopAfterRecord: bool = False

    def session_prefix(self) -> str:
        """<savePath>/<timestamp>[_<fileName>]: shared by every file of one recording session (recorder.cpp:77-82, octprozapp.cpp:296)"""
        return os.path.join(self.savePath, self.timestamp + (("_" + self.fileName) if self.fileName else ""))

    def for_processed_data(self, q) -> "RecordingParams":
        """what Processing::slot_enableRecording hands to the processed-data recorder (processing.cpp:243-249): float32 buffers of
        N/2 x A x B, or half the raw buffer's bytes (the converted output keeps the raw container, half as many samples)"""
        import copy as _copy
        r = _copy.copy(self)
        if self.saveAs32bitFloat:
            r.bufferSizeInBytes = (int(q.samplesPerLine) // 2) * int(q.ascansPerBscan) * int(q.bscansPerBuffer) * 4
        else:
            r.bufferSizeInBytes = self.bufferSizeInBytes // 2
        return r

    def save_meta(self, settings_file: str) -> str | None:
        """the recording's meta file is a copy of the settings INI (octprozapp.cpp:294-298); returns its path"""
        if not self.saveMetaData:
            return None
        import shutil
        dst = self.session_prefix() + "_meta.txt"
        shutil.copyfile(settings_file, dst)
        return dst


class Recorder:
    """octproz/src/recorder.cpp.  Two ways in:
      Recorder(path, buffers_to_record) + record(buf)      -- N buffers appended to one headerless file (the format the Virtual OCT
                                                              System replays), recorder.cpp:99-152 in its plainest form;
      Recorder("raw" | "processed") + slot_init(RecordingParams) + slot_record(buf, ..., currentBufferNr) + slot_abortRecording()
                                                           -- the reference's recording session: file name
                                                              <savePath>/<timestamp>[_<fileName>]_<name>.raw (:77-82), optional start
                                                              at the first buffer of a volume (:116-119), capture in memory and one
                                                              write when the last buffer has arrived or on abort (:124-131, :52-62).
    Messages the reference emits as signals are collected in `messages` as (kind, text)."""

    def __init__(self, path_or_name: str, buffers_to_record: int | None = None):
        self.messages: list[tuple[str, str]] = []
        self.on_recording_done = None        # signal recordingDone (recorder.h)
        self.on_ready_to_record = None       # signal readyToRecord(bool): Processing switches float streaming on it (processing.cpp:258)
        self.recorded = self.recordedBuffers = 0
        self.recordingEnabled = self.recordingFinished = self.isRecording = self.initialized = False
        self._chunks: list[bytes] = []
        self._f = None
        if buffers_to_record is not None:     # plain form: stream to the file as the buffers come
            self.name, self.path, self.buffers_to_record = "", path_or_name, buffers_to_record
            self._f = open(path_or_name, "wb")
        else:
            self.name, self.path, self.buffers_to_record = path_or_name, "", 0
            self.currRecParams = RecordingParams()

    # ---- plain form
    def record(self, buf: np.ndarray) -> bool:
        if self._f is None or self.recorded >= self.buffers_to_record:
            return False
        self._f.write(np.ascontiguousarray(buf).tobytes())
        self.recorded += 1
        if self.recorded == self.buffers_to_record:
            self._f.close()
        return True

    # ---- the reference's session
    def _emit_ready(self, ready: bool) -> None:
        if self.on_ready_to_record:
            self.on_ready_to_record(ready)

    def slot_init(self, rec_params: RecordingParams) -> bool:
        self.currRecParams = rec_params
        if not rec_params.savePath or not os.path.isdir(rec_params.savePath):
            self.messages.append(("error", "Recording not initialized: save path is empty or invalid."))
            self._uninit()
            return False
        self.path = rec_params.session_prefix() + "_" + self.name + ".raw"
        self._chunks, self.recordedBuffers = [], 0
        self.initialized, self.recordingFinished, self.recordingEnabled, self.isRecording = True, False, True, False
        self._emit_ready(True)
        self.messages.append(("info", "Recording initialized..."))
        return True

    def _uninit(self) -> None:
        self._chunks = []
        self.initialized, self.recordingFinished, self.recordedBuffers = False, True, 0
        self._emit_ready(False)
        if self.on_recording_done:
            self.on_recording_done()

    def slot_record(self, buffer, bitDepth=0, samplesPerLine=0, linesPerFrame=0, framesPerBuffer=0, buffersPerVolume=0, currentBufferNr=0) -> None:
        if not self.recordingEnabled:
            return
        if not self.initialized:
            self.messages.append(("error", "Recording not possible. Record buffer not initialized."))
            return
        if self.currRecParams.startWithFirstBuffer and not self.isRecording and currentBufferNr != 0:
            return                                                     # wait for the first buffer of a volume
        self.isRecording = True
        raw = np.ascontiguousarray(buffer).view(np.uint8).reshape(-1)
        want = int(self.currRecParams.bufferSizeInBytes)
        if raw.size < want:
            raise ValueError(f"buffer holds {raw.size} bytes, the recording expects {want} per buffer")
        self._chunks.append(raw[:want].tobytes())                      # the reference copies bufferSizeInBytes, whatever the buffer holds
        self.recordedBuffers += 1
        if self.recordedBuffers >= self.currRecParams.buffersToRecord:
            self.recordingEnabled = self.isRecording = False
            self._save_to_disk()
            self._uninit()

    def slot_abortRecording(self) -> None:
        if self.recordingEnabled and not self.recordingFinished:
            self.messages.append(("error", "Recording aborted!"))
            self.recordingEnabled = False
            self._save_to_disk()                                       # what has been captured so far is kept
            self._uninit()

    def _save_to_disk(self) -> None:
        if not self.initialized:
            self.messages.append(("error", "Save recording to disk not possible. Record buffer not initialized."))
            return
        try:
            with open(self.path, "wb") as f:
                for c in self._chunks:
                    f.write(c)
        except OSError:
            self.messages.append(("error", "Recording failed! Could not write file to disk."))
            return
        self.messages.append(("info", f"Captured buffers: {self.recordedBuffers}/{self.currRecParams.buffersToRecord}"))
        self.messages.append(("info", "Data written to disk! " + self.path))


class Processing:
    """Processing::slot_start (processing.cpp:136-229) without Qt: `pipeline` is an OctPipeline (or a stand-in with the same
    initializeCuda / octCudaPipeline / sync / cleanupCuda methods)."""

    def __init__(self, pipeline, oct_params):
        self.pipeline, self.octParams = pipeline, oct_params
        self.stats = {}
        self.on_raw_data = None          # signal rawData(ptr, bitDepth, N, A, B, buffersPerVolume, currentBufferNr) (processing.h:110)
        self.processed_buffers = 0
        self.rawRecorder, self.processedRecorder = Recorder("raw"), Recorder("processed")      # processing.cpp:49,60
        self.messages: list[tuple[str, str]] = []
        self._curr_nr = 0
        self._stream_bufs = None
        self._memorized = None

    # ---- recording (processing.cpp:231-266, octprozapp.cpp:225-299, 408-422)
    def slot_enableRecording(self, rec_params: RecordingParams, settings_file: str | None = None) -> None:
        """start a recording session: raw buffers as Processing sees them and / or the processed buffers the pipeline streams to the
        host (converted containers, or float32 with saveAs32bitFloat); the meta file is a copy of `settings_file`.  Call before
        slot_start or while it runs."""
        q = self.octParams
        if rec_params.recordRaw:
            if self.rawRecorder.recordingEnabled:
                self.messages.append(("error", "Recording of raw data is already running."))
            else:
                self.rawRecorder.slot_init(rec_params)
        if rec_params.recordProcessed:
            if self.processedRecorder.recordingEnabled:
                self.messages.append(("error", "Recording of processed data is already running."))
            else:
                # slot_prepareGpu2HostForProcessedRecording: every buffer is streamed while the recording runs, settings restored after
                self._memorized = (q.streamToHost, q.streamingBuffersToSkip, q.saveAs32bitFloat)
                q.streamToHost, q.streamingBuffersToSkip, q.saveAs32bitFloat = True, 0, bool(rec_params.saveAs32bitFloat)
                self.processedRecorder.on_recording_done = self._processed_recording_done
                self.processedRecorder.slot_init(rec_params.for_processed_data(q))
                self._streaming_wanted = True
        if settings_file is not None:
            rec_params.save_meta(settings_file)

    def _processed_recording_done(self) -> None:
        q = self.octParams
        if self._memorized is not None:                               # slot_resetGpu2HostSettings
            q.streamToHost, q.streamingBuffersToSkip, q.saveAs32bitFloat = self._memorized
            self._memorized = None

    def _enable_streaming(self) -> None:
        """enableGpu2HostStreaming / enableFloatGpu2HostStreaming (processing.cpp:316-362): two host buffers registered with the
        pipeline; its callbacks (Gpu2HostNotifier, gpu2hostnotifier.h:47-49) hand every delivered buffer to the processed recorder"""
        q, rec = self.octParams, self.processedRecorder
        as_float = bool(q.saveAs32bitFloat)
        nbytes = int(rec.currRecParams.bufferSizeInBytes)
        bufs = AcquisitionBuffer(); bufs.allocateMemory(2, nbytes)
        by_addr = {int(b.ctypes.data): b for b in bufs.bufferArray}

        def deliver(ptr):
            b = by_addr.get(int(ptr or 0))
            if b is not None:
                nr = getattr(self.pipeline, "current_buffer_nr", None)        # params->currentBufferNr of the streaming call (cuda_code.cu:1602)
                rec.slot_record(b, q.bitDepth, q.samplesPerLine // 2, q.ascansPerBscan, q.bscansPerBuffer, q.buffersPerVolume,
                                nr() if callable(nr) else self._curr_nr)
        if as_float:
            self.pipeline.cuda_registerFloatStreamingBuffers(bufs.bufferArray[0], bufs.bufferArray[1], nbytes)
            self.pipeline.set_callbacks(float_streaming=deliver)
        else:
            self.pipeline.cuda_registerStreamingBuffers(bufs.bufferArray[0], bufs.bufferArray[1], nbytes)
            self.pipeline.set_callbacks(streaming=deliver)
        self._stream_bufs = (bufs, as_float)
        self._streaming_wanted = False

    def _disable_streaming(self) -> None:
        if self._stream_bufs is None:
            return
        bufs, as_float = self._stream_bufs
        self.pipeline.sync()
        if as_float:
            self.pipeline.cuda_unregisterFloatStreamingBuffers()
        else:
            self.pipeline.cuda_unregisterStreamingBuffers()
        bufs.releaseMemory()
        self._stream_bufs = None

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
                self._curr_nr = curr_nr
                if getattr(self, "_streaming_wanted", False):
                    self._enable_streaming()
                if self.on_raw_data:
                    self.on_raw_data(buf.bufferArray[pos], q.bitDepth, q.samplesPerLine, q.ascansPerBscan, q.bscansPerBuffer, q.buffersPerVolume, curr_nr)
                self.rawRecorder.slot_record(buf.bufferArray[pos], q.bitDepth, q.samplesPerLine, q.ascansPerBscan, q.bscansPerBuffer,
                                             q.buffersPerVolume, curr_nr)                                          # connect(rawData, rawRecorder) :52
                self.pipeline.octCudaPipeline(buf.bufferArray[pos])  # :187
                buf.bufferReadyArray[pos] = False                     # :191
                n += 1
                if max_buffers is not None and n >= max_buffers:
                    system.stopAcquisition()
            else:
                time.sleep(0)
        self.pipeline.sync()
        dt = time.perf_counter() - t0
        self._disable_streaming()
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
