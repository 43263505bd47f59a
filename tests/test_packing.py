"""12-bit packed raw format (octproz_b200/packing.py): layout and round trip on CPU."""
import numpy as np
import pytest

from octproz_b200.packing import pack12, unpack12


def test_round_trip_and_bit_layout():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 4096, (3, 5, 64)).astype(np.uint16)
    p = pack12(a)
    assert p.dtype == np.uint8 and p.shape == (3, 5, 96)
    assert np.array_equal(unpack12(p), a)
    # GenICam Mono12p: sample k of a line occupies bits [12k, 12k+12) of the line's little-endian bit string
    x = (np.arange(8, dtype=np.uint16) * 0x111) & 0xFFF
    bits = int.from_bytes(pack12(x).tobytes(), "little")
    assert all(((bits >> (12 * k)) & 0xFFF) == int(x[k]) for k in range(8))
    assert pack12(np.array([0xABC, 0x123], np.uint16)).tolist() == [0xBC, 0x3A, 0x12]
    # edge cases: empty, full scale, rejects
    assert pack12(np.zeros((2, 0), np.uint16)).shape == (2, 0)
    assert np.array_equal(unpack12(pack12(np.full(32, 4095, np.uint16))), np.full(32, 4095, np.uint16))
    with pytest.raises(ValueError):
        pack12(np.zeros(3, np.uint16))
    with pytest.raises(ValueError):
        pack12(np.array([4096, 0], np.uint16))
    with pytest.raises(ValueError):
        unpack12(np.zeros(4, np.uint8))
