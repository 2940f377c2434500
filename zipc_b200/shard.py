"""Host-side partitioning for the multi-GPU path: members / slices are independent units,
one process per GPU, no collective on the data path; only tiny per-unit results are gathered and checksums
of adjacent slices are merged with zipc_b200_crc32_combine / zipc_b200_adler32_combine."""
from __future__ import annotations

import heapq
from typing import Sequence

from . import _lib


def partition_lpt(sizes: Sequence[int], world: int) -> list[list[int]]:
    """Longest-processing-time-first: members sorted by size, each to the least loaded rank."""
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    parts: list[list[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda k: -int(sizes[k])):
        load, r = heapq.heappop(heap)
        parts[r].append(i)
        heapq.heappush(heap, (load + int(sizes[i]), r))
    return parts


def slice_bounds(n: int, world: int, align: int = 512) -> list[tuple[int, int]]:
    """Contiguous slices of one buffer, boundaries aligned (the last slice takes the remainder)."""
    per = (n // world) // align * align
    out = []
    for r in range(world):
        lo = r * per
        hi = n if r == world - 1 else (r + 1) * per
        out.append((lo, hi))
    return out


def combine_crc32(parts: Sequence[tuple[int, int]]) -> int:
    """parts = [(crc, length)] of adjacent slices in order -> CRC-32 of the concatenation."""
    L = _lib.lib()
    acc = None
    for crc, n in parts:
        acc = crc if acc is None else L.zipc_b200_crc32_combine(acc, crc, n)
    return 0 if acc is None else acc


def combine_adler32(parts: Sequence[tuple[int, int]]) -> int:
    """RFC 1950 Adler-32 of the concatenation of adjacent slices."""
    L = _lib.lib()
    acc = None
    for ad, n in parts:
        acc = ad if acc is None else L.zipc_b200_adler32_combine(acc, ad, n)
    return 1 if acc is None else acc
