#!/usr/bin/env python
"""bench.py -- headline measurement of the zipc hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload crc32|inflate|deflate] [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic input.
  crc32   (default, BASELINE.json configs[1]): CRC-32 of a 1 GiB rand-v1 buffer (test_crc_speed shape)
  inflate (configs[2]): batch inflate + CRC-32 of 10,000 deflate members of 4-256 KiB text-v1
  deflate (configs[3]): batch deflate (level default) + CRC-32 of the same 10,000 members
`value` is device-resident throughput in GB/s of UNCOMPRESSED bytes (CUDA events on the library's own
stream); `e2e` is the same metric through the C-ABI call with host buffers, H2D and D2H inside the timed
region.  Inputs are larger than L2 (>= 1 GiB per step), so no L2 flush is needed between iterations.
With N > 1 (torchrun) every rank runs the same per-GPU workload on its own GPU (weak scaling: members /
buffers are independent units, no data-path collective); rank 0 prints the aggregate.  The default run
also reports the two secondary workloads under "also" (smaller step counts) unless --no-also is given.

--impl reference times the reference's algorithm on the host cores: the OCaml reference cannot be built in
this image (no OCaml toolchain), so it is the C restatement in oracle/ ("port").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GiB = 1 << 30
METRIC = {"crc32": "crc32_GBps_uncompressed", "inflate": "inflate_GBps_uncompressed", "deflate": "deflate_GBps_uncompressed"}
WORKLOAD_DESC = {
    "crc32": "C2: CRC-32 over a 1 GiB rand-v1(seed=2) buffer per GPU (test_crc_speed shape)",
    "inflate": "C3: batch inflate + CRC-32 of 10,000 deflate members (text-v1, 4-256 KiB, 1.33 GB) per GPU",
    "deflate": "C4: batch deflate (level default) + CRC-32 of 10,000 members (text-v1, 4-256 KiB, 1.33 GB) per GPU",
}


def ncu_traffic(which):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu capture of this command (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return int(json.load(f)[which]["traffic_bytes"])
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
def make_members(count: int, seed0: int = 1000):
    from zipc_b200 import synth
    sizes = synth.member_sizes(count, seed=3)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        datas = list(ex.map(lambda a: synth.text_v1(seed0 + a[0], int(a[1])), enumerate(sizes)))
    return datas


class Harness:
    def __init__(self, device: int):
        import torch
        from zipc_b200 import zipc_deflate as zd
        self.torch, self.zd = torch, zd
        torch.cuda.set_device(device)
        self.ctx = zd.Context(device)
        self.L = self.ctx.L
        self.stream = torch.cuda.ExternalStream(self.ctx.stream)
        self.device = device

    def pinned(self, arr: np.ndarray) -> np.ndarray:
        p = C.c_void_p()
        assert self.L.zipc_b200_host_alloc(max(arr.size, 1), C.byref(p)) == 0
        out = np.ctypeslib.as_array((C.c_uint8 * max(arr.size, 1)).from_address(p.value))[:arr.size]
        out[:] = arr
        return out

    def timed(self, fn, steps, warmup):
        """K steps bracketed by CUDA events on the library stream -> total ms."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for _ in range(steps):
            fn()
        b.record(self.stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    def kernel_ms(self, fn, steps):
        """average duration of the dominant kernel alone (events recorded by the library around it)"""
        self.L.zipc_b200_ctx_profile(self.ctx.h, 1)
        ts = []
        for _ in range(steps):
            fn()
            ts.append(self.L.zipc_b200_ctx_kernel_ms(self.ctx.h))
        self.L.zipc_b200_ctx_profile(self.ctx.h, 0)
        ts = [t for t in ts if t > 0]
        return float(np.mean(ts)) if ts else None


def run_crc32(h: Harness, steps, warmup, rank):
    from zipc_b200 import synth
    import zlib
    n = GiB
    host = h.pinned(synth.rand_v1(2 + rank, n))
    d = h.torch.from_numpy(host).to(f"cuda:{h.device}")
    dcrc = h.torch.zeros(4, dtype=h.torch.int32, device=f"cuda:{h.device}")
    fn = lambda: h.L.zipc_b200_crc32_dev_async(h.ctx.h, d.data_ptr(), n, dcrc.data_ptr())
    l0 = h.ctx.launches
    total_ms = h.timed(fn, steps, warmup)
    launches = (h.ctx.launches - l0) // (steps + warmup) * steps
    got = int(dcrc[0].item()) & 0xFFFFFFFF
    kms = h.kernel_ms(fn, min(steps, 10))
    # end to end: host (pinned) buffer in, 4 bytes out, every step
    out = C.c_uint32()
    e2e_fn = lambda: h.L.zipc_b200_crc32(h.ctx.h, host.ctypes.data, n, C.byref(out))
    for _ in range(2):
        e2e_fn()
    t0 = time.perf_counter()
    esteps = max(3, min(steps, 5))
    for _ in range(esteps):
        e2e_fn()
    e2e_s = (time.perf_counter() - t0) / esteps
    check = zlib.crc32(host[: 64 << 20])  # cheap spot check of the generator + full check of the result below
    assert out.value == got, (hex(out.value), hex(got))
    # Adler-32 over the same resident buffer (synchronous call: chunk kernel + device fold + 4-byte read back)
    adler = {}
    for name, mode in (("ref_compat", 0), ("rfc1950", 1)):
        aout = C.c_uint32()
        afn = lambda: h.L.zipc_b200_adler32_dev(h.ctx.h, d.data_ptr(), n, mode, C.byref(aout))
        for _ in range(3):
            assert afn() == 0
        t0 = time.perf_counter()
        for _ in range(10):
            afn()
        dt = (time.perf_counter() - t0) / 10
        akms = h.kernel_ms(afn, 5)
        adler[name] = {"value": "%08x" % aout.value, "GBps_per_call": round(n / dt / 1e9, 1), "chunk_kernel_ms": round(akms, 4) if akms else None}
    assert int(adler["rfc1950"]["value"], 16) == zlib.adler32(host), adler
    return dict(units=n, total_ms=total_ms, launches=launches, kernel_ms=kms, algo_bytes=n, e2e_s=e2e_s,
                h2d=n, d2h=4, result=got, host=host, extra={"crc32": "%08x" % got, "spot": "%08x" % check, "adler32": adler})


def _pack_device(h: Harness, items):
    """concatenate (16-byte aligned) into one pinned host buffer + device copy; returns offsets"""
    offs = np.zeros(len(items), dtype=np.uint64)
    lens = np.array([len(x) for x in items], dtype=np.uint64)
    t = 0
    for i, x in enumerate(items):
        offs[i] = t
        t += (len(x) + 15) & ~15
    host = np.zeros(t + 64, dtype=np.uint8)
    for i, x in enumerate(items):
        host[int(offs[i]):int(offs[i]) + len(x)] = np.frombuffer(x, dtype=np.uint8) if not isinstance(x, np.ndarray) else x
    hp = h.pinned(host)
    return hp, h.torch.from_numpy(hp).to(f"cuda:{h.device}"), offs, lens


def run_inflate_foreign(h: Harness, count=10000):
    """C3 with a FOREIGN encoder: the same members compressed by zlib -6 on the host (longer matches, ~16k-symbol
    blocks, occasional stored / fixed blocks) and inflated on the GPU, device-resident."""
    import zlib
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    datas = make_members(count, 1000)

    def comp(d):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        return np.frombuffer(c.compress(d.tobytes()) + c.flush(), dtype=np.uint8)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        streams = list(ex.map(comp, datas))
    crcs = np.array([zlib.crc32(d.tobytes()) for d in datas[:200]], dtype=np.uint32)
    n = len(datas)
    U = int(sum(d.size for d in datas)); Cb = int(sum(x.size for x in streams))
    _, dcs, coff, clen = _pack_device(h, streams)
    soff = np.concatenate([[0], np.cumsum([(d.size + 15) & ~15 for d in datas])[:-1]]).astype(np.uint64)
    slen = np.array([d.size for d in datas], dtype=np.uint64)
    ddst = h.torch.empty(int(soff[-1] + slen[-1]) + 64, dtype=h.torch.uint8, device=f"cuda:{h.device}")
    dl = np.zeros(n, dtype=np.uint64); ck = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)

    def fn():
        rc = h.L.zipc_b200_inflate_batch_dev(h.ctx.h, 2, 0, n, dcs.data_ptr(), P(coff, C.c_size_t), P(clen, C.c_size_t),
                                             ddst.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t), P(dl, C.c_size_t),
                                             P(ck, C.c_uint32), P(st, C.c_int))
        assert rc == 0, rc
    for _ in range(3):
        fn()
    h.torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    h.torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    assert (st == 0).all() and (dl == slen).all() and (ck[:200] == crcs).all(), "parity lost on zlib-made streams"
    return {"workload": "C3 members compressed by zlib -6 on the host (foreign encoder), batch inflate + CRC-32, device-resident",
            "metric": "inflate_GBps_uncompressed", "value": round(U / dt / 1e9, 2), "unit": "GB/s", "ratio": round(Cb / U, 4), "members": n}


def run_archive(h: Harness, count=10000):
    """ZIP archive layer end to end (host buffers): Zipc.File.deflate_of_binary_string x n + Zipc.to_binary_string
    in one call, then Zipc.of_binary_string + File.to_binary_string x n (CRC-32 checked) in two."""
    import io
    import zipfile
    from zipc_b200 import _lib
    datas = make_members(count, 1000)
    n = len(datas)
    U = int(sum(d.size for d in datas))
    off = np.concatenate([[0], np.cumsum([d.size for d in datas])[:-1]]).astype(np.int64)
    src = h.pinned(np.concatenate(datas))
    names = [b"dir%03d/member%05d.txt" % (i % 97, i) for i in range(n)]
    paths = (C.c_void_p * n)(*[C.cast(C.c_char_p(x), C.c_void_p).value for x in names])  # `names` keeps the bytes alive
    plen = np.array([len(x) for x in names], dtype=np.uint32)
    ptrs = (C.c_void_p * n)(*[src.ctypes.data + int(o) for o in off])
    slen = np.array([d.size for d in datas], dtype=np.uint64)
    out = h.pinned(np.zeros(U // 2 + 256 * n + 65536, dtype=np.uint8))
    olen = C.c_size_t()
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))

    def create():
        rc = h.L.zipc_b200_zip_deflate_archive(h.ctx.h, 2, n, paths, P(plen, C.c_uint32), ptrs, P(slen, C.c_size_t), None, None, None,
                                               out.ctypes.data, out.size, C.byref(olen))
        assert rc == 0, rc
    create()
    t0 = time.perf_counter(); create(); tc = time.perf_counter() - t0
    alen = olen.value
    # an independent reader sees the same members
    zf = zipfile.ZipFile(io.BytesIO(bytes(out[:alen])))
    assert len(zf.namelist()) == n
    for i in (0, n // 2, n - 1):
        assert zf.read(names[i].decode()) == datas[i].tobytes()
    arena = h.pinned(np.zeros(U + 16 * n + 4096, dtype=np.uint8))
    need = C.c_size_t(); doff = np.zeros(n, dtype=np.uint64); dlen = np.zeros(n, dtype=np.uint64)
    found = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)

    def extract():
        ms = C.POINTER(_lib.Member)(); cnt = C.c_size_t()
        rc = h.L.zipc_b200_zip_parse(out.ctypes.data, alen, C.byref(ms), C.byref(cnt))
        assert rc == 0 and cnt.value == n, (rc, cnt.value)
        rc = h.L.zipc_b200_zip_extract_batch(h.ctx.h, ms, n, arena.ctypes.data, arena.size, C.byref(need), P(doff, C.c_size_t),
                                             P(dlen, C.c_size_t), P(found, C.c_uint32), P(st, C.c_int))
        h.L.zipc_b200_free(ms)
        assert rc == 0 and (st == 0).all(), rc
    extract()
    t0 = time.perf_counter(); extract(); tx = time.perf_counter() - t0
    assert int(dlen.sum()) == U
    return {"workload": "ZIP archive of the C3/C4 members (10k files, default level): create = zipc_b200_zip_deflate_archive, "
                        "extract = zipc_b200_zip_parse + zipc_b200_zip_extract_batch (CRC-32 checked); pinned host buffers",
            "members": n, "uncompressed_bytes": U, "archive_bytes": int(alen),
            "create_e2e_GBps": round(U / tc / 1e9, 3), "extract_e2e_GBps": round(U / tx / 1e9, 3),
            "checked": "python zipfile lists all members and reads three of them back"}


def run_codec(h: Harness, which, steps, warmup, rank, count=10000, level="default"):
    from zipc_b200 import _lib
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    datas = make_members(count, 1000 + rank * count)
    U = int(sum(d.size for d in datas))
    lvl = {"fast": 1, "default": 2, "best": 3}[level]
    hsrc, dsrc, soff, slen = _pack_device(h, datas)
    n = len(datas)
    # compressed streams made by the GPU encoder (validated against the oracle in tests/)
    res = h.ctx.deflate_batch(datas, level, _lib.CK_CRC32)
    assert all(r[0] == 0 for r in res)
    streams = [r[1] for r in res]
    crcs = np.array([r[2] for r in res], dtype=np.uint32)
    Cb = int(sum(s.size for s in streams))
    dl = np.zeros(n, dtype=np.uint64); ck = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)
    if which == "inflate":
        hcs, dcs, coff, clen = _pack_device(h, streams)
        ddst = h.torch.empty(int(soff[-1] + ((slen[-1] + 15) & ~np.uint64(15))) + 64, dtype=h.torch.uint8, device=f"cuda:{h.device}")

        def fn():
            rc = h.L.zipc_b200_inflate_batch_dev(h.ctx.h, 2, 0, n, dcs.data_ptr(), P(coff, C.c_size_t), P(clen, C.c_size_t),
                                                 ddst.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t), P(dl, C.c_size_t),
                                                 P(ck, C.c_uint32), P(st, C.c_int))
            assert rc == 0, rc
        # end to end: compressed members inside one pinned host buffer (an in-memory archive), pinned output arena
        e_ptrs = (C.c_void_p * n)(*[hcs.ctypes.data + int(o) for o in coff])
        e_arena = h.pinned(np.zeros(int(sum((int(x) + 15) & ~15 for x in slen)) + 64, dtype=np.uint8))
        e_need = C.c_size_t(); e_off = np.zeros(n, dtype=np.uint64)

        def e2e_fn():
            rc = h.L.zipc_b200_inflate_batch(h.ctx.h, 2, 0, n, e_ptrs, P(clen, C.c_size_t), P(slen, C.c_size_t), e_arena.ctypes.data,
                                             e_arena.size, C.byref(e_need), P(e_off, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
            assert rc == 0 and (st == 0).all(), rc
        h2d, d2h = Cb, U
    else:
        cap = np.array([h.L.zipc_b200_deflate_bound(int(x)) + 15 & ~15 for x in slen], dtype=np.uint64)
        doff = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint64)
        ddst = h.torch.empty(int(cap.sum()) + 64, dtype=h.torch.uint8, device=f"cuda:{h.device}")

        def fn():
            rc = h.L.zipc_b200_deflate_batch_dev(h.ctx.h, lvl, 2, 0, n, dsrc.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t),
                                                 ddst.data_ptr(), P(doff, C.c_size_t), P(cap, C.c_size_t), P(dl, C.c_size_t),
                                                 P(ck, C.c_uint32), P(st, C.c_int))
            assert rc == 0, rc

        e_ptrs = (C.c_void_p * n)(*[hsrc.ctypes.data + int(o) for o in soff])
        e_arena = h.pinned(np.zeros(U // 2 + 16 * n + 4096, dtype=np.uint8))
        e_need = C.c_size_t(); e_off = np.zeros(n, dtype=np.uint64)

        def e2e_fn():
            rc = h.L.zipc_b200_deflate_batch(h.ctx.h, lvl, 2, 0, n, e_ptrs, P(slen, C.c_size_t), e_arena.ctypes.data, e_arena.size,
                                             C.byref(e_need), P(e_off, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
            assert rc == 0 and (st == 0).all(), rc
        h2d, d2h = U, Cb
    l0 = h.ctx.launches
    # the _dev entry points are synchronous (they return per-member results), so wall clock == device time
    for _ in range(warmup):
        fn()
    h.torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    h.torch.cuda.synchronize()
    total_ms = (time.perf_counter() - t0) * 1e3
    launches = (h.ctx.launches - l0) // (steps + warmup) * steps
    assert (st == 0).all() and (ck == crcs).all(), "parity lost in bench"
    kms = h.kernel_ms(fn, min(steps, 5))
    e2e_fn()
    t0 = time.perf_counter()
    esteps = max(2, min(steps, 3))
    for _ in range(esteps):
        e2e_fn()
    e2e_s = (time.perf_counter() - t0) / esteps
    return dict(units=U, total_ms=total_ms, launches=launches, kernel_ms=kms, algo_bytes=U + Cb, e2e_s=e2e_s,
                h2d=h2d, d2h=d2h, datas=datas, streams=streams,
                extra={"members": n, "uncompressed_bytes": U, "compressed_bytes": Cb, "ratio": round(Cb / U, 4), "level": level})


# ---------------------------------------------------------------------------------------------------
# CPU baselines (oracle = C restatement of the reference; kind "port")
# ---------------------------------------------------------------------------------------------------
def cpu_crc32(host: np.ndarray, budget_s=10.0):
    from oracle import zipc_oracle as zo
    L = zo.lib()
    n = host.size
    t0 = time.perf_counter()
    reps = 0
    while True:
        L.zo_crc32(C.cast(host.ctypes.data, C.c_char_p), n)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 8:
            break
    dt = time.perf_counter() - t0
    import zlib
    t1 = time.perf_counter(); zlib.crc32(host[: 256 << 20]); tz = time.perf_counter() - t1   # a familiar yardstick
    return {"value": round(n * reps / dt / 1e9, 3), "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": f"{reps} x full 1 GiB buffer, Crc_32.string restated in C (oracle/zipc_oracle.c), 1 thread: the reference hashes one string on one core",
            "zlib_crc32_yardstick_GBps": round((256 << 20) / tz / 1e9, 3)}


def cpu_codec(which, datas, streams, level="default", budget_s=12.0):
    """one member per core from a shared queue (the reference's natural multi-core use)"""
    from oracle import zipc_oracle as zo
    zo.lib()
    cores = os.cpu_count() or 1
    items = list(zip(datas, streams))
    # bounded sample: as many members as fit the budget, estimated from a probe
    probe = items[:cores]
    def work(it):
        d, s = it
        if which == "inflate":
            out, crc = zo.inflate_and_crc_32(s.tobytes(), d.size)
            return d.size
        zo.crc_32_and_deflate(d.tobytes(), level)
        return d.size
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done = sum(ex.map(work, probe))
    rate = done / max(time.perf_counter() - t0, 1e-6)
    target = int(rate * budget_s)
    sample, acc = [], 0
    for it in items:
        if acc >= target:
            break
        sample.append(it); acc += it[0].size
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        done = sum(ex.map(work, sample))
    dt = time.perf_counter() - t0
    return {"value": round(done / dt / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"{len(sample)} of {len(items)} members ({done/1e6:.0f} MB uncompressed), one member per thread over {cores} threads, "
                      f"{'inflate_and_crc_32' if which == 'inflate' else 'crc_32_and_deflate level ' + level} restated in C (oracle/)"}


# ---------------------------------------------------------------------------------------------------
def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from zipc_b200 import synth
    which = args.workload
    steps, warmup = args.steps, args.warmup
    if which == "crc32":
        # weak scaling: N GPUs hash N independent 1 GiB buffers, so the reference hashes N buffers too, one
        # string per core (its only form of parallelism)
        nbuf = max(1, min(args.gpus, os.cpu_count() or 1))
        hosts = [synth.rand_v1(2 + r, GiB) for r in range(nbuf)]
        from oracle import zipc_oracle as zo
        L = zo.lib()
        one = lambda hbuf: L.zo_crc32(C.cast(hbuf.ctypes.data, C.c_char_p), hbuf.size)
        def step():
            with ThreadPoolExecutor(max_workers=nbuf) as ex:
                list(ex.map(one, hosts))
        for _ in range(min(warmup, 1)):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
        value, cores = GiB * nbuf * steps / dt / 1e9, nbuf
        sample = f"each step = Crc_32.string (C restatement) over {nbuf} x 1 GiB buffer(s), one string per thread (one string = one core in the reference)"
    else:
        count = 600
        datas = make_members(count)
        from oracle import zipc_oracle as zo
        cores = os.cpu_count() or 1
        if which == "inflate":
            with ThreadPoolExecutor(max_workers=cores) as ex:
                streams = list(ex.map(lambda d: zo.deflate(d.tobytes(), "default"), datas))
            work = lambda it: len(zo.inflate_and_crc_32(it[1], it[0].size)[0])
        else:
            streams = [None] * count
            work = lambda it: (zo.crc_32_and_deflate(it[0].tobytes(), "default"), it[0].size)[1]
        items = list(zip(datas, streams))
        def step():
            with ThreadPoolExecutor(max_workers=cores) as ex:
                return sum(ex.map(work, items))
        for _ in range(min(warmup, 1)):
            step()
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            done += step()
        dt = time.perf_counter() - t0
        value = done / dt / 1e9
        sample = f"each step = {count} members of the C3/C4 distribution ({sum(d.size for d in datas)/1e6:.0f} MB), one member per thread over {cores} threads"
    line = {"impl": "reference", "metric": METRIC[which], "value": round(value, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[which], "note": "reference is OCaml (no toolchain in this image): timed through its C restatement oracle/zipc_oracle.c"},
            "cpu_baseline": {"value": round(value, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", choices=["crc32", "inflate", "deflate"], default="crc32")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads in the default run")
    ap.add_argument("--members", type=int, default=10000)
    ap.add_argument("--level", choices=["fast", "default", "best"], default="default", help="deflate level of the codec workloads")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.workload == "crc32" else 5
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
    h = Harness(local)
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(which, steps, warmup):
        barrier()
        with ClockSampler(local) as cs:
            r = run_crc32(h, steps, warmup, rank) if which == "crc32" else run_codec(h, which, steps, warmup, rank, args.members, args.level)
        barrier()
        r["clocks"] = cs.summary()
        # max over ranks of the device time and of the end-to-end time; units summed over ranks
        vals = torch.tensor([r["total_ms"], r["e2e_s"]], dtype=torch.float64, device=f"cuda:{local}")
        units = torch.tensor([float(r["units"])], dtype=torch.float64, device=f"cuda:{local}")
        if dist is not None:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
            dist.all_reduce(units, op=dist.ReduceOp.SUM)
        r["total_ms_max"], r["e2e_s_max"], r["units_all"] = float(vals[0]), float(vals[1]), float(units[0])
        return r

    which = args.workload
    r = run(which, args.steps, args.warmup)
    value = r["units_all"] * args.steps / (r["total_ms_max"] / 1e3) / 1e9
    e2e = r["units_all"] / r["e2e_s_max"] / 1e9
    roof = None
    if r["kernel_ms"]:
        ach = r["algo_bytes"] / (r["kernel_ms"] / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": ncu_traffic(which), "kernel": {"crc32": "crc32_tiles_kernel", "inflate": "inflate_kernel<false>", "deflate": "deflate_kernel"}[which],
                "kernel_ms": round(r["kernel_ms"], 4), "algorithmic_bytes": r["algo_bytes"], "peak_source": peak_src}
    line = {"metric": METRIC[which], "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(r["total_ms_max"] / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[which].replace("level default", "level " + args.level), "l2": "inputs larger than L2 (>= 1 GiB per step), no flush needed",
                       "per_gpu_bytes": int(r["units"]), **r["extra"]},
            "clocks": r["clocks"],
            "e2e": {"value": round(e2e, 3), "unit": "GB/s", "h2d_bytes_per_step": int(r["h2d"]), "d2h_bytes_per_step": int(r["d2h"]),
                    "note": "C-ABI call with pinned host buffers (codec members lie in one pinned buffer, like an in-memory archive); H2D + kernels + D2H per step"},
            "gpu_launches": int(r["launches"]), "roofline": roof}
    if rank == 0:
        if which == "crc32":
            line["cpu_baseline"] = cpu_crc32(r["host"]) if world == 1 else None
        else:
            line["cpu_baseline"] = cpu_codec(which, r["datas"], r["streams"]) if world == 1 else None
    if which == "crc32" and not args.no_also and world == 1:
        also = {}
        del r
        for w in ("inflate", "deflate"):
            try:
                rr = run(w, 3, 3)
                ach = rr["algo_bytes"] / (rr["kernel_ms"] / 1e3) / 1e9 if rr["kernel_ms"] else None
                also[w] = {"metric": METRIC[w], "value": round(rr["units_all"] * 3 / (rr["total_ms_max"] / 1e3) / 1e9, 2), "unit": "GB/s",
                           "e2e": round(rr["units_all"] / rr["e2e_s_max"] / 1e9, 3), "kernel_ms": rr["kernel_ms"],
                           "roofline_frac": round(ach / peak, 4) if ach else None, "workload": WORKLOAD_DESC[w], **rr["extra"],
                           "cpu_baseline": cpu_codec(w, rr["datas"], rr["streams"], budget_s=8.0)}
                del rr
            except Exception as e:  # never lose the headline line to a secondary workload
                also[w] = {"error": repr(e)}
        try:
            also["inflate_foreign_zlib6"] = run_inflate_foreign(h)
        except Exception as e:
            also["inflate_foreign_zlib6"] = {"error": repr(e)}
        try:
            also["archive"] = run_archive(h)
        except Exception as e:
            also["archive"] = {"error": repr(e)}
        try:  # C1: one 64 MiB text-v1 stream, segment-independent deflate + indexed inflate, pinned host buffers (e2e)
            from zipc_b200 import synth
            L = h.L
            data = h.pinned(synth.text_v1(1, 64 << 20))
            cbuf = h.pinned(np.zeros(data.size + (data.size >> 3) + 65536, dtype=np.uint8))
            obuf = h.pinned(np.zeros(data.size, dtype=np.uint8))
            res = {}
            for seg in (16 << 10, 64 << 10, 256 << 10):
                nmax = -(-data.size // seg) + 1
                index = np.zeros((nmax + 1, 2), dtype=np.uint64)
                ip = index.ctypes.data_as(C.POINTER(C.c_uint64))
                n, nseg, crc, crc2, st, olen = C.c_size_t(), C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_int(), C.c_size_t()
                dfn = lambda: L.zipc_b200_deflate_segmented(h.ctx.h, 2, data.ctypes.data, data.size, seg, 1, cbuf.ctypes.data, cbuf.size,
                                                            C.byref(n), ip, nmax + 1, C.byref(nseg), C.byref(crc))
                ifn = lambda: L.zipc_b200_inflate_segmented(h.ctx.h, cbuf.ctypes.data, n.value, ip, nseg.value, obuf.ctypes.data, obuf.size,
                                                            C.byref(olen), C.byref(crc2), C.byref(st))
                assert dfn() == 0
                t0 = time.perf_counter()
                for _ in range(3): dfn()
                td = (time.perf_counter() - t0) / 3
                assert ifn() == 0 and st.value == 0
                t0 = time.perf_counter()
                for _ in range(3): ifn()
                ti = (time.perf_counter() - t0) / 3
                assert crc2.value == crc.value and olen.value == data.size and bytes(obuf[:4096]) == bytes(data[:4096])
                res["seg_%dk" % (seg >> 10)] = {"deflate_e2e_GBps": round(data.size / td / 1e9, 3), "inflate_e2e_GBps": round(data.size / ti / 1e9, 3),
                                               "ratio": round(n.value / data.size, 4), "segments": int(nseg.value)}
            also["stream_c1"] = {"workload": "C1: 64 MiB text-v1(seed=1) as one RFC 1951 stream of independent segments + index; CRC-32 fused both ways; pinned host buffers in and out", **res}
        except Exception as e:
            also["stream_c1"] = {"error": repr(e)}
        line["also"] = also
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
