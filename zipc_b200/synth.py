"""Synthetic workload generators of SURVEY.md section 8d (integer only, host side, in the library)."""
from __future__ import annotations

import numpy as np

from . import _lib


def text_v1(seed: int, n: int) -> np.ndarray:
    """Zipf-word text: same vocabulary for every seed, different word sequence."""
    out = np.empty(n, dtype=np.uint8)
    if n:
        _lib.lib().zipc_b200_synth_text(seed, out.ctypes.data, n)
    return out


def rand_v1(seed: int, n: int) -> np.ndarray:
    """Raw splitmix64 output, little endian."""
    out = np.empty(n, dtype=np.uint8)
    if n:
        _lib.lib().zipc_b200_synth_rand(seed, out.ctypes.data, n)
    return out


def member_sizes(count: int, seed: int = 3) -> np.ndarray:
    """C3/C4 member sizes: 4096 + r % 258049 (uniform 4 KiB .. 256 KiB), splitmix64 stream."""
    raw = rand_v1(seed, 8 * count).view(np.uint64)
    return (4096 + raw % np.uint64(258049)).astype(np.int64)
