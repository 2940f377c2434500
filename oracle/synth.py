"""The synthetic workloads of SURVEY.md 8d for the reference arm of bench.py: the same generators as
zipc_b200.synth (one source file, zipc_b200/csrc/synth.cc), loaded from oracle/libzipc_synth.so so that the
reference process never maps the product library.  TEST / MEASUREMENT INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libzipc_synth.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "..", "zipc_b200", "csrc", "synth.cc")
        if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
            subprocess.run(["make", "-C", _HERE, "-s", "libzipc_synth.so"], check=True)
        L = C.CDLL(_SO)
        for f in (L.zipc_b200_synth_text, L.zipc_b200_synth_rand):
            f.restype, f.argtypes = None, [C.c_uint64, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def text_v1(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    if n:
        lib().zipc_b200_synth_text(seed, out.ctypes.data, n)
    return out


def rand_v1(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    if n:
        lib().zipc_b200_synth_rand(seed, out.ctypes.data, n)
    return out


def member_sizes(count: int, seed: int = 3) -> np.ndarray:
    raw = rand_v1(seed, 8 * count).view(np.uint64)
    return (4096 + raw % np.uint64(258049)).astype(np.int64)
