"""octproz_b200 -- B200-native OCT raw -> B-scan pipeline behind OCTproZ's processing API.

The product is liboctb200.so (CUDA, sm_100a, C ABI in include/octb200.h); this package is the thin
host side: the ctypes binding, the mirror of the reference's parameter object, the `Processing`-style
driver, the headless Virtual-OCT-System replay and the multi-GPU sharding helper.
"""
from . import _lib  # noqa: F401
from .params import OctAlgorithmParameters, benchmark_params  # noqa: F401
from .pipeline import OctPipeline  # noqa: F401

__all__ = ["OctAlgorithmParameters", "OctPipeline", "benchmark_params"]
