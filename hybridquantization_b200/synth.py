"""Synthetic inputs of SURVEY.md 8(d): SplitMix64-seeded RGB images and java.util.Random palettes.

numpy only (no oracle, no GPU).  tests/test_synth.py checks these against the oracle's C versions.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x48510000  # + config number (C1..C5)
PALETTE_SEED = 77760

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(seed: int, first: int, count: int) -> np.ndarray:
    """outputs first .. first+count-1 (0-based) of SplitMix64 seeded with `seed`."""
    with np.errstate(over="ignore"):
        i = np.arange(first + 1, first + 1 + count, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_image(width: int, height: int, seed: int, smooth: bool = False) -> np.ndarray:
    """uint8 [height, width, 3].  uniform: byte j = byte (j%8) of SplitMix64 output j//8 + 2.
    smooth: per-channel bilinear blend of four corner bytes (outputs 0, 1) + (byte % 17) - 8, clamped."""
    nbytes = width * height * 3
    words = _splitmix64(seed, 2, (nbytes + 7) // 8)
    img = words.view(np.uint8)[:nbytes].copy() if words.dtype.byteorder in ("<", "=", "|") else None
    img = img.reshape(height, width, 3)
    if not smooth:
        return img
    o = _splitmix64(seed, 0, 2).view(np.uint8)  # 16 bytes: corners are bytes 0..11
    wd, hd = max(width - 1, 1), max(height - 1, 1)
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    out = np.empty_like(img)
    for c in range(3):
        c00, c01, c10, c11 = (int(o[c * 4 + q]) for q in range(4))
        top = c00 * (wd - x) + c01 * x
        bot = c10 * (wd - x) + c11 * x
        base = (top * (hd - y) + bot * y) // (wd * hd)
        v = base + (img[:, :, c].astype(np.int64) % 17) - 8
        out[:, :, c] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def synth_image_rows(width: int, height: int, seed: int, r0: int, r1: int) -> np.ndarray:
    """Rows [r0, r1) of the UNIFORM synth_image(width, height, seed) without generating the rest
    (a rank's shard of a large image)."""
    b0, b1 = r0 * width * 3, r1 * width * 3
    w0, w1 = b0 // 8, (b1 + 7) // 8
    words = _splitmix64(seed, 2 + w0, max(w1 - w0, 0))
    return words.view(np.uint8)[b0 - 8 * w0: b1 - 8 * w0].copy().reshape(r1 - r0, width, 3)


class _JavaRandomNp:
    """java.util.Random in pure Python ints (host-side test data only)."""

    def __init__(self, seed: int):
        self.s = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def next_float_block(self, count: int) -> np.ndarray:
        out = np.empty(count, np.float32)
        s = self.s
        mask = (1 << 48) - 1
        for i in range(count):
            s = (s * 0x5DEECE66D + 0xB) & mask
            out[i] = np.float32(s >> 24) / np.float32(1 << 24)
        self.s = s
        return out


def synth_palettes(B: int, K: int, seed: int = PALETTE_SEED) -> np.ndarray:
    """float32 [B, K, 4]: B palettes drawn consecutively in SWASA.generateRandomColors order
    (SWASA.java:40-52) from java.util.Random(seed)."""
    r = _JavaRandomNp(seed)
    pal = np.zeros((B, K, 4), np.float32)
    pal[:, :, :3] = r.next_float_block(B * K * 3).reshape(B, K, 3)
    return pal
