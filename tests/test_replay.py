"""Headless Virtual-OCT-System replay + Processing loop (octproz_b200/acquisition.py): the double-buffer handshake,
file format and rate statistics of the reference (virtualoctsystem.cpp:163-224, processing.cpp:136-229), on CPU with a
stand-in pipeline, and end to end on the GPU."""
import copy

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
