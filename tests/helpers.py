import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def from_bits(hexlist, shape=None):
    a = np.array([int(h, 16) for h in hexlist], np.uint32).view(np.float32)
    return a.reshape(shape) if shape else a


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def fbits(x: float) -> int:
    return int(np.float32(x).view(np.uint32))


def idx_crc(idx, n):
    return int(np.bitwise_xor.reduce((idx.astype(np.uint64) + 1) * (np.arange(n, dtype=np.uint64) * 2654435761 % 4294967291)))
