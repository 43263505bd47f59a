"""The reference's settings-file vocabulary ([processing] keys of sidebar.h:58-94, [Virtual%20OCT%20System] keys of
virtualoctsystemsettingsdialog.h:27-38, [streaming]): the Python reader (octproz_b200/params.py from_ini) and the C++ reader
(include/octb200_host.hpp OctAlgorithmParameters::fromIni) on a file with the published benchmark values, against each other and
against benchmark_params(); where the reference tree is present, on its own checked-in benchmark file as well."""
import json
import os
import subprocess

import numpy as np
import pytest

from octproz_b200 import OctAlgorithmParameters, benchmark_params
from tests.test_host_mirror import build

# the published benchmark settings (performance/v180/20250504_performance_v180_gtx1080/20250504_octproz_settings.ini), typed in here
PROCESSING = {"addend": "0", "bitshift": "false", "coeff": "1", "dispersion_compensation": "true", "dispersion_compensation_d0": "0",
              "dispersion_compensation_d1": "97", "dispersion_compensation_d2": "-96.625", "dispersion_compensation_d3": "-0.375",
              "fixed_pattern_removal": "true", "fixed_pattern_removal_continuously": "false", "fixed_pattern_removal_bscans": "1",
              "flip_bscans": "false", "log": "true", "max": "100", "min": "-30", "resampling": "true", "resampling_c0": "0.535239",
              "resampling_c1": "871.817574", "resampling_c2": "-170.633784", "resampling_c3": "97.249716", "resampling_interpolation": "1",
              "sinusoidal_scan_correction": "false", "window_center_position": "0.5", "window_fill_factor": "0.95", "window_type": "0",
              "windowing": "true", "background_removal": "false", "background_removal_window_size": "8",
              "post_processing_background_removal": "false", "post_processing_background_removal_offset": "0",
              "post_processing_background_removal_weight": "1"}
VOS = {"bit_depth": "12", "buffers_from_file": "2", "buffers_per_volume": "1", "depth": "256", "file_path": "C:/test_data_raw.raw", "height": "512",
       "wait_time": "0", "width": "1024", "bscan_offset": "0", "sync_with_processing": "true"}
REFERENCE_FILE = "/root/reference/performance/v180/20250504_performance_v180_gtx1080/20250504_octproz_settings.ini"
FIELDS = ("bitshift", "bscanFlip", "signalLogScaling", "sinusoidalScanCorrection", "signalGrayscaleMin", "signalGrayscaleMax", "signalMultiplicator",
          "signalAddend", "backgroundRemoval", "rollingAverageWindowSize", "resampling", "resamplingInterpolation", "dispersionCompensation", "windowing",
          "window", "windowFillFactor", "windowCenter", "fixedPatternNoiseRemoval", "continuousFixedPatternNoiseDetermination", "bscansForNoiseDetermination",
          "postProcessBackgroundRemoval", "postProcessBackgroundWeight", "postProcessBackgroundOffset", "bitDepth", "samplesPerLine", "ascansPerBscan",
          "bscansPerBuffer", "buffersPerVolume")


def write_ini(path, streaming=True):
    with open(path, "w") as f:
        f.write("[General]\ntimestamp=0\n\n[processing]\n" + "".join(f"{k}={v}\n" for k, v in PROCESSING.items()))
        f.write(f"\n[streaming]\nstreaming_enabled={'true' if streaming else 'false'}\nstreaming_skip=2\n")
        f.write("\n[Virtual%20OCT%20System]\n" + "".join(f"{k}={v}\n" for k, v in VOS.items()))


# benchmark_params() leaves the (disabled) rolling-average window at the constructor default; the file says 8
SAME_AS_BENCHMARK = tuple(f for f in FIELDS if f != "rollingAverageWindowSize") + ("c0", "c1", "c2", "c3", "d0", "d1", "d2", "d3")


def test_python_reader_gives_the_benchmark_parameters(tmp_path):
    path = str(tmp_path / "settings.ini")
    write_ini(path)
    q, b = OctAlgorithmParameters.from_ini(path), benchmark_params()
    for name in SAME_AS_BENCHMARK:
        assert getattr(q, name) == getattr(b, name), name
    assert q.streamToHost and q.streamingBuffersToSkip == 2
    assert q.rollingAverageWindowSize == 8 and not q.backgroundRemoval
    q.update_all_curves(); b.update_all_curves()
    assert np.array_equal(q.resampleCurve, b.resampleCurve) and np.array_equal(q.windowCurve, b.windowCurve)


def cpp_parse(tmp_path, ini):
    exe = str(tmp_path / "host_mirror_test")
    if not os.path.exists(exe):
        build("tests/host/host_mirror_test.cpp", exe)
    r = subprocess.run([exe, str(tmp_path / "replay.raw"), str(tmp_path / "window.f32"), ini], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("INI ")]
    assert line, r.stdout
    return json.loads(line[0][4:])


def check_cpp_against_python(cpp, q):
    for name in FIELDS:
        want = getattr(q, name)
        got = cpp[name]
        if isinstance(want, float):
            assert got == pytest.approx(float(np.float32(want)), rel=1e-7, abs=0), name       # the C ABI carries floats
        else:
            assert got == int(want), name
    assert cpp["c"] == pytest.approx([float(np.float32(x)) for x in (q.c0, q.c1, q.c2, q.c3)], rel=1e-7)
    assert cpp["d"] == pytest.approx([float(np.float32(x)) for x in (q.d0, q.d1, q.d2, q.d3)], rel=1e-7)
    assert cpp["streamToHost"] == int(q.streamToHost) and cpp["streamingBuffersToSkip"] == q.streamingBuffersToSkip


def test_cpp_reader_agrees_with_the_python_reader(tmp_path):
    path = str(tmp_path / "settings.ini")
    write_ini(path, streaming=False)
    cpp = cpp_parse(tmp_path, path)
    check_cpp_against_python(cpp, OctAlgorithmParameters.from_ini(path))
    assert cpp["buffersFromFile"] == 2 and cpp["bscanOffset"] == 0 and cpp["syncWithProcessing"] == 1 and cpp["streamToHost"] == 0


@pytest.mark.skipif(not os.path.exists(REFERENCE_FILE), reason="reference tree not present (GPU box)")
def test_both_readers_on_the_reference_benchmark_file(tmp_path):
    q, b = OctAlgorithmParameters.from_ini(REFERENCE_FILE), benchmark_params()
    for name in SAME_AS_BENCHMARK:
        assert getattr(q, name) == getattr(b, name), name
    assert q.streamToHost                      # the published benchmark runs with streaming to the host enabled
    check_cpp_against_python(cpp_parse(tmp_path, REFERENCE_FILE), q)
