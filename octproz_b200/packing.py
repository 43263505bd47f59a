"""12-bit packed raw data (GenICam PFNC "Mono12p"): two samples per three bytes, little-endian bit order -- sample k of a line
occupies bits [12k, 12k+12) of the line's bit string.  The reference takes 12-bit data only in 16-bit containers
(docs/docs/faq.md); `OctPipeline(input_packing=PACK_12P)` accepts this format directly and moves a quarter fewer bytes over PCIe.
Host helpers for producers / tests (numpy only)."""
from __future__ import annotations

import numpy as np


def pack12(samples: np.ndarray) -> np.ndarray:
    """u16 containers (values < 4096, even count along the last axis) -> uint8 array with 3/2 bytes per sample"""
    a = np.ascontiguousarray(samples, np.uint16)
    if a.shape[-1] % 2:
        raise ValueError("need an even number of samples per line")
    if a.size and int(a.max()) > 4095:
        raise ValueError("12-bit samples expected")
    lo = a[..., 0::2].astype(np.uint32); hi = a[..., 1::2].astype(np.uint32)
    w = lo | (hi << 12)                                   # 24 bits per pair
    out = np.empty(a.shape[:-1] + (a.shape[-1] // 2, 3), np.uint8)
    out[..., 0] = w & 0xFF; out[..., 1] = (w >> 8) & 0xFF; out[..., 2] = (w >> 16) & 0xFF
    return out.reshape(a.shape[:-1] + (a.shape[-1] * 3 // 2,))


def unpack12(packed: np.ndarray) -> np.ndarray:
    """inverse of pack12 -> uint16 containers"""
    b = np.ascontiguousarray(packed, np.uint8)
    if b.shape[-1] % 3:
        raise ValueError("packed line length must be a multiple of 3 bytes")
    t = b.reshape(b.shape[:-1] + (b.shape[-1] // 3, 3)).astype(np.uint32)
    w = t[..., 0] | (t[..., 1] << 8) | (t[..., 2] << 16)
    out = np.empty(b.shape[:-1] + (b.shape[-1] // 3, 2), np.uint16)
    out[..., 0] = w & 0xFFF; out[..., 1] = (w >> 12) & 0xFFF
    return out.reshape(b.shape[:-1] + (b.shape[-1] * 2 // 3,))
