"""Synthetic inputs restated from the reference's own generators so measurements are
comparable with its benches: gradient_rgb (tests/support/synthetic.rs:74-85,
benches/comparison.rs:32-43) and the LCG noise (synthetic.rs:183-197)."""
from __future__ import annotations

import numpy as np


def gradient_rgb(width: int, height: int) -> np.ndarray:
    x = np.arange(width, dtype=np.uint64)[None, :]
    y = np.arange(height, dtype=np.uint64)[:, None]
    out = np.empty((height, width, 3), np.uint8)
    out[..., 0] = (x * 255 // max(width, 1)).astype(np.uint8)
    out[..., 1] = np.broadcast_to((y * 255 // max(height, 1)).astype(np.uint8), (height, width))
    out[..., 2] = ((x + y) * 127 // max(width + height, 1)).astype(np.uint8)
    return out.reshape(-1)


def noise(width: int, height: int, channels: int = 3, seed: int = 42) -> np.ndarray:
    """state = state*1103515245 + 12345 (u32 wrap); byte = state >> 16; one draw per byte."""
    n = int(width) * int(height) * int(channels)
    a = np.uint32(1103515245)
    c = np.uint32(12345)
    with np.errstate(over="ignore"):
        an = np.cumprod(np.full(n, a, np.uint32), dtype=np.uint32)          # a^(k+1)
        # c * (a^k + ... + 1): prefix sums of a^j, j = 0..k
        geo = np.cumsum(np.concatenate([np.ones(1, np.uint32), an[:-1]]), dtype=np.uint32)
        state = an * np.uint32(seed & 0xFFFFFFFF) + geo * c
    return (state >> np.uint32(16)).astype(np.uint8)
