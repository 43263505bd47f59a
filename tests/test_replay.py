"""Headless Virtual-OCT-System replay + Processing loop (octproz_b200/acquisition.py): the double-buffer handshake,
file format and rate statistics of the reference (virtualoctsystem.cpp:163-224, processing.cpp:136-229), on CPU with a
stand-in pipeline, and end to end on the GPU."""
import copy
import os

import numpy as np
import pytest

from octproz_b200 import benchmark_params, synth
from octproz_b200.acquisition import AcquisitionBuffer, Processing, Recorder, VirtualOCTSystem, read_raw_file, replay, write_raw_file


class FakePipeline:
    def __init__(self, ok=True):
        self.ok, self.seen, self.init_args = ok, [], None

    def initializeCuda(self, h1, h2, q):
        self.init_args = (h1, h2, q)
        return self.ok

    def octCudaPipeline(self, h):
        self.seen.append((h.ctypes.data, bytes(h[:8])))

    def sync(self):
        pass


def test_acquisition_buffer_alignment_and_flags():
    b = AcquisitionBuffer()
    assert b.allocateMemory(2, 1000)
    assert all(a.ctypes.data % 128 == 0 and a.nbytes == 1000 for a in b.bufferArray)      # acquisitionbuffer.cpp:65
    assert b.bufferReadyArray == [False, False] and b.currIndex == -1
    b.releaseMemory()
    assert b.bufferArray == []


def test_raw_file_roundtrip_and_bscan_offset(tmp_path):
    vol = synth.make_volume(64, 4, 6, 12)
    p = str(tmp_path / "v.raw")
    write_raw_file(p, vol)
    assert (tmp_path / "v.raw").stat().st_size == vol.size * 2                            # headerless
    assert np.array_equal(read_raw_file(p, 12, 64, 4, 6), vol)
    assert np.array_equal(read_raw_file(p, 12, 64, 4, 2, bscan_offset=3), vol[3:5])       # virtualoctsystem.cpp:167


@pytest.mark.parametrize("buffers_from_file", [1, 2])
def test_replay_alternates_the_two_buffers(tmp_path, buffers_from_file):
    n, a, b = 64, 4, 3
    vol = synth.make_volume(n, a, 2 * b, 12)
    p = str(tmp_path / "v.raw")
    write_raw_file(p, vol)
    q = benchmark_params(n, a, b); q.buffersPerVolume = 2
    fake = FakePipeline()
    raw_signals = []
    vos_kw = dict(buffers_from_file=buffers_from_file)
    from octproz_b200.acquisition import VirtualOCTSystem as V
    proc = replay(p, q, fake, buffers=7, **vos_kw)
    assert proc.processed_buffers == 7 and len(fake.seen) == 7
    ptrs = [s[0] for s in fake.seen]
    assert len(set(ptrs)) == 2 and all(ptrs[i] != ptrs[i + 1] for i in range(6))          # strict alternation of the two host buffers
    first = bytes(vol[:b].reshape(-1)[:4].tobytes()); second = bytes(vol[b:2 * b].reshape(-1)[:4].tobytes())
    contents = {s[1] for s in fake.seen}
    assert contents == ({first} if buffers_from_file == 1 else {first, second})             # :175-179
    st = proc.stats
    assert st["ascans_per_s"] == pytest.approx(st["buffers_per_s"] * b * a) and st["volumes_per_s"] == pytest.approx(st["buffers_per_s"] / 2)
    assert fake.init_args[0].ctypes.data in ptrs and fake.init_args[1].ctypes.data in ptrs


def test_replay_streams_a_large_file_buffer_by_buffer(tmp_path):
    """buffers_from_file > 2 (acqcuisitionSimulationLargeFile, virtualoctsystem.cpp:226-290): successive buffers of the file, starting
    at bscan_offset, rewinding after buffers_from_file buffers"""
    n, a, b = 64, 4, 2
    vol = synth.make_volume(n, a, 6 * b, 12)
    p = str(tmp_path / "big.raw")
    write_raw_file(p, vol)
    q = benchmark_params(n, a, b)
    fake = FakePipeline()
    proc = replay(p, q, fake, buffers=9, buffers_from_file=4, bscan_offset=b)
    assert proc.processed_buffers == 9
    # consecutive buffers of the file, cyclic; the start-up handshake (Processing raises, then clears all ready flags around
    # initializeCuda, processing.cpp:124-134) may discard the first delivered buffer, in the reference as well
    heads = [bytes(vol[b * (1 + k): b * (2 + k)].reshape(-1)[:4].tobytes()) for k in range(4)]
    off = heads.index(fake.seen[0][1])
    assert [s[1] for s in fake.seen] == [heads[(i + off) % 4] for i in range(9)]
    ptrs = [s[0] for s in fake.seen]
    assert len(set(ptrs)) == 2 and all(ptrs[i] != ptrs[i + 1] for i in range(8))


def test_failed_initialisation_stops_the_acquisition(tmp_path):
    vol = synth.make_volume(64, 4, 2, 12)
    p = str(tmp_path / "v.raw"); write_raw_file(p, vol)
    q = benchmark_params(64, 4, 2)
    proc = replay(p, q, FakePipeline(ok=False), buffers=3)
    assert proc.processed_buffers == 0                                                      # initializationFailed -> stop (processing.cpp:152-156)


def test_recorder_writes_headerless_buffers(tmp_path):
    r = Recorder(str(tmp_path / "rec.raw"), 2)
    x = np.arange(10, dtype=np.uint16)
    assert r.record(x) and r.record(x + 1) and not r.record(x + 2)
    assert np.array_equal(np.fromfile(str(tmp_path / "rec.raw"), np.uint16), np.concatenate([x, x + 1]))


@pytest.mark.gpu
def test_replay_through_the_cuda_pipeline(tmp_path):
    from octproz_b200 import OctPipeline
    from oracle import oracle as orc
    from tests.util import assert_parity
    n, a, b = 1024, 32, 4
    q = benchmark_params(n, a, b); q.fixedPatternNoiseRemoval = False; q.update_all_curves()
    vol = synth.make_volume(n, a, 2 * b, 12, resample=q.resampleCurve, dispersion=q.dispersionCurve)
    p = str(tmp_path / "v.raw"); write_raw_file(p, vol)
    pipe = OctPipeline()
    proc = replay(p, copy.deepcopy(q), pipe, buffers=5)
    assert proc.processed_buffers == 5 and proc.stats["ascans_per_s"] > 0
    out = pipe.copy_output(0)
    pipe.cleanupCuda()
    from tests.util import parity_report
    # the last processed buffer is one of the two halves of the file (which one depends on where the handshake started)
    reports = [parity_report(out, orc.process(q, vol[i * b:(i + 1) * b])[0], q) for i in range(2)]
    assert min(r["frac_outside"] for r in reports) <= 1e-4, reports


def test_packed12_file_replay_delivers_packed_buffers(tmp_path):
    """extension: a raw file of 12-bit packed samples is replayed with 3/2 bytes per sample"""
    from octproz_b200.packing import pack12, unpack12
    n, a, b = 64, 4, 2
    vol = (np.arange(2 * b * a * n, dtype=np.uint32) * 7 % 4096).astype(np.uint16).reshape(2 * b, a, n)
    path = str(tmp_path / "packed.raw")
    pack12(vol).tofile(path)
    vos = VirtualOCTSystem(path, 12, n, a, b, packed12=True, sync_with_processing=False)
    seen = []
    vos.on_acquisition_started = lambda s: (seen.extend(unpack12(np.array(s.buffer.bufferArray[i][: b * a * n * 3 // 2])) for i in (0, 1)), s.stopAcquisition())
    vos.startAcquisition()
    assert len(seen) == 2
    assert np.array_equal(seen[0].reshape(b, a, n), vol[:b]) and np.array_equal(seen[1].reshape(b, a, n), vol[b:])


# ---------------------------------------------------------------- recording sessions (recorder.cpp, processing.cpp:231-266)
class FakeStreamingPipeline(FakePipeline):
    """stand-in with the streaming half of the kernels.h surface: the "processed" buffer is the raw buffer's first half + 1 (u16) or
    the same as float32, delivered through the registered host buffers and the callback, alternating between the two buffers"""

    def __init__(self):
        super().__init__()
        self.stream = self.fstream = None
        self.cb = self.fcb = None
        self.calls = 0
        self.nr, self.q = 0, None

    def initializeCuda(self, h1, h2, q):
        self.q = q
        self.nr = q.buffersPerVolume - 1
        return super().initializeCuda(h1, h2, q)

    def cuda_registerStreamingBuffers(self, h1, h2, nbytes):
        self.stream = (h1, h2, nbytes)

    def cuda_unregisterStreamingBuffers(self):
        self.stream = None

    def cuda_registerFloatStreamingBuffers(self, h1, h2, nbytes):
        self.fstream = (h1, h2, nbytes)

    def cuda_unregisterFloatStreamingBuffers(self):
        self.fstream = None

    def set_callbacks(self, streaming=None, float_streaming=None, background=None):
        self.cb, self.fcb = streaming, float_streaming

    def current_buffer_nr(self):
        return self.nr

    @staticmethod
    def processed(raw_u16):
        return (raw_u16[: raw_u16.size // 2] + 1).astype(np.uint16)

    def octCudaPipeline(self, h):
        super().octCudaPipeline(h)
        self.nr = (self.nr + 1) % self.q.buffersPerVolume
        raw = h.view(np.uint16).reshape(-1)
        if self.q.streamToHost and self.stream and not self.q.saveAs32bitFloat:
            dst = self.stream[self.calls & 1]
            dst.view(np.uint16)[:] = self.processed(raw)
            self.cb(dst.ctypes.data)
        if self.q.saveAs32bitFloat and self.fstream:
            dst = self.fstream[self.calls & 1]
            dst.view(np.float32)[:] = self.processed(raw).astype(np.float32)
            self.fcb(dst.ctypes.data)
        self.calls += 1


def test_recorder_session_file_name_first_buffer_rule_and_abort(tmp_path):
    from octproz_b200.acquisition import RecordingParams
    rp = RecordingParams(timestamp="20261017_101500", fileName="phantom", savePath=str(tmp_path), bufferSizeInBytes=20, buffersToRecord=3,
                         startWithFirstBuffer=True)
    done = []
    r = Recorder("raw"); r.on_recording_done = lambda: done.append(1)
    x = np.arange(10, dtype=np.uint16)
    r.slot_record(x)                                              # not initialised, not enabled: ignored (recorder.cpp:107)
    assert r.slot_init(rp) and r.path == str(tmp_path / "20261017_101500_phantom_raw.raw")
    r.slot_record(x + 100, currentBufferNr=1)                     # waits for the first buffer of a volume (:116)
    assert r.recordedBuffers == 0 and not r.isRecording
    for i, nr in enumerate((0, 1, 0, 1)):                         # the fourth is past buffersToRecord
        r.slot_record(x + i, currentBufferNr=nr)
    assert done == [1] and r.recordingFinished and not r.recordingEnabled
    assert np.array_equal(np.fromfile(r.path, np.uint16), np.concatenate([x, x + 1, x + 2]))
    assert ("info", "Captured buffers: 3/3") in r.messages
    # no user file name -> no extra underscore; abort keeps what was captured
    rp2 = RecordingParams(timestamp="t", savePath=str(tmp_path), bufferSizeInBytes=20, buffersToRecord=5)
    r2 = Recorder("processed"); assert r2.slot_init(rp2) and os.path.basename(r2.path) == "t_processed.raw"
    r2.slot_record(x); r2.slot_record(x + 7)
    r2.slot_abortRecording()
    assert np.array_equal(np.fromfile(r2.path, np.uint16), np.concatenate([x, x + 7])) and ("error", "Recording aborted!") in r2.messages
    r2.slot_abortRecording()                                       # second abort: nothing to do
    # invalid save path
    r3 = Recorder("raw")
    assert not r3.slot_init(RecordingParams(timestamp="t", savePath=str(tmp_path / "missing"), bufferSizeInBytes=20, buffersToRecord=1))
    assert r3.messages[0][0] == "error" and not r3.recordingEnabled


@pytest.mark.parametrize("as_float", [False, True])
def test_processing_records_raw_and_processed_sessions(tmp_path, as_float):
    from octproz_b200.acquisition import RecordingParams
    n, a, b = 64, 4, 3
    vol = synth.make_volume(n, a, 2 * b, 12)
    p = str(tmp_path / "v.raw"); write_raw_file(p, vol)
    ini = tmp_path / "settings.ini"; ini.write_text("[processing]\nbitshift=false\n")
    q = benchmark_params(n, a, b); q.buffersPerVolume = 2
    q.streamToHost, q.streamingBuffersToSkip = False, 3
    vos = VirtualOCTSystem(p, q.bitDepth, n, a, b, q.buffersPerVolume)
    fake = FakeStreamingPipeline()
    proc = Processing(fake, q)
    rp = RecordingParams(timestamp="ts", fileName="", savePath=str(tmp_path), bufferSizeInBytes=n * a * b * 2, buffersToRecord=4,
                         startWithFirstBuffer=True, recordRaw=True, recordProcessed=True, saveMetaData=True, saveAs32bitFloat=as_float)
    proc.slot_enableRecording(rp, settings_file=str(ini))
    assert q.streamToHost and q.streamingBuffersToSkip == 0                          # every buffer is streamed while the recording runs
    proc.slot_enableRecording(rp)                                                     # a second request while one is running
    assert ("error", "Recording of raw data is already running.") in proc.messages and ("error", "Recording of processed data is already running.") in proc.messages
    import threading
    started = threading.Event(); vos.on_acquisition_started = lambda s: started.set()
    t = threading.Thread(target=vos.startAcquisition, daemon=True); t.start()
    assert started.wait(30)
    assert proc.slot_start(vos, max_buffers=9)
    vos.stopAcquisition(); t.join(30)
    assert (tmp_path / "ts_meta.txt").read_text() == ini.read_text()
    halves = [vol[:b].reshape(-1), vol[b:].reshape(-1)]
    raw = np.fromfile(str(tmp_path / "ts_raw.raw"), np.uint16).reshape(4, -1)
    # four consecutive buffers; the recording started with a first buffer of a volume (currentBufferNr 0) and the file's two halves alternate
    first = 0 if np.array_equal(raw[0], halves[0]) else 1
    assert all(np.array_equal(raw[i], halves[(first + i) & 1]) for i in range(4))
    if as_float:
        got = np.fromfile(str(tmp_path / "ts_processed.raw"), np.float32).reshape(4, -1)
        assert got.shape[1] == (n // 2) * a * b
        want = [FakeStreamingPipeline.processed(h).astype(np.float32) for h in halves]
    else:
        got = np.fromfile(str(tmp_path / "ts_processed.raw"), np.uint16).reshape(4, -1)
        assert got.shape[1] == (n // 2) * a * b                                     # half the raw buffer's bytes (processing.cpp:248)
        want = [FakeStreamingPipeline.processed(h) for h in halves]
    f0 = 0 if np.array_equal(got[0], want[0]) else 1
    assert all(np.array_equal(got[i], want[(f0 + i) & 1]) for i in range(4))
    # settings restored once the processed recording is complete (slot_resetGpu2HostSettings), streaming buffers released at the end
    assert not q.streamToHost and q.streamingBuffersToSkip == 3 and not q.saveAs32bitFloat
    assert fake.stream is None and fake.fstream is None
    assert proc.rawRecorder.recordingFinished and proc.processedRecorder.recordingFinished
