"""Scratch GPU probe (not part of the product): quick timings of the kernels through the C ABI.
Usage: python tools/gpu_probe.py [crc] [adler] [inflate] [deflate]"""
import ctypes as C
import sys
import time
import zlib

import numpy as np
import torch

sys.path.insert(0, ".")
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

what = set(sys.argv[1:]) or {"crc", "adler", "inflate"}
ctx = zd.Context(0)
L = ctx.L
stream = torch.cuda.ExternalStream(ctx.stream)
print(torch.cuda.get_device_name(0), "SMs", torch.cuda.get_device_properties(0).multi_processor_count)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(stream)
        fn()
        b.record(stream)
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[0], ts[len(ts) // 2]


if "crc" in what:
    for n in (64 << 20, 1 << 30):
        host = synth.rand_v1(2, n)
        d = torch.from_numpy(host).cuda()
        dcrc = torch.zeros(4, dtype=torch.int32, device="cuda")
        fn = lambda: L.zipc_b200_crc32_dev_async(ctx.h, d.data_ptr(), n, dcrc.data_ptr())
        best, med = timed(fn)
        got = int(dcrc[0].item()) & 0xFFFFFFFF
        ok = got == zlib.crc32(host)
        print(f"crc32 n={n>>20} MiB best {best:.3f} ms med {med:.3f} ms -> {n/best/1e6:.1f} GB/s (med {n/med/1e6:.1f}) ok={ok}")
        del d

if "adler" in what:
    n = 1 << 30
    host = synth.rand_v1(2, n)
    d = torch.from_numpy(host).cuda()
    out = C.c_uint32()
    for mode in (0, 1):
        t0 = time.time()
        for _ in range(5):
            L.zipc_b200_adler32_dev(ctx.h, d.data_ptr(), n, mode, C.byref(out))
        dt = (time.time() - t0) / 5
        print(f"adler32 mode={mode} wall {dt*1e3:.3f} ms -> {n/dt/1e9:.1f} GB/s value {out.value:08x}")
    print("zlib adler", hex(zlib.adler32(host)))
    del d

if "inflate" in what:
    nmem = int([a for a in sys.argv if a.startswith("n=")][0][2:]) if any(a.startswith("n=") for a in sys.argv) else 2000
    sizes = synth.member_sizes(nmem, seed=7)
    t0 = time.time()
    datas = [synth.text_v1(1000 + i, int(s)).tobytes() for i, s in enumerate(sizes)]
    streams = []
    for x in datas:
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        streams.append(c.compress(x) + c.flush())
    print(f"built {nmem} members in {time.time()-t0:.1f}s: U={sum(map(len,datas))/1e6:.1f} MB C={sum(map(len,streams))/1e6:.1f} MB")
    U = sum(map(len, datas))
    src_off = np.zeros(nmem, dtype=np.uint64); dst_off = np.zeros(nmem, dtype=np.uint64)
    so = do = 0
    for i in range(nmem):
        src_off[i] = so; so += (len(streams[i]) + 15) & ~15
        dst_off[i] = do; do += (len(datas[i]) + 15) & ~15
    hsrc = np.zeros(so, dtype=np.uint8)
    for i in range(nmem):
        hsrc[src_off[i]:src_off[i] + len(streams[i])] = np.frombuffer(streams[i], dtype=np.uint8)
    dsrc = torch.from_numpy(hsrc).cuda()
    ddst = torch.zeros(do + 64, dtype=torch.uint8, device="cuda")
    sl = np.array([len(s) for s in streams], dtype=np.uint64); mo = np.array([len(x) for x in datas], dtype=np.uint64)
    dl = np.zeros(nmem, dtype=np.uint64); ck = np.zeros(nmem, dtype=np.uint32); st = np.zeros(nmem, dtype=np.int32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    for ckind in (0, 2):
        def fn():
            rc = L.zipc_b200_inflate_batch_dev(ctx.h, ckind, 0, nmem, dsrc.data_ptr(), P(src_off, C.c_size_t), P(sl, C.c_size_t),
                                               ddst.data_ptr(), P(dst_off, C.c_size_t), P(mo, C.c_size_t), P(dl, C.c_size_t),
                                               P(ck, C.c_uint32), P(st, C.c_int))
            assert rc == 0, rc
        fn()
        t0 = time.time()
        for _ in range(3):
            fn()
        dt = (time.time() - t0) / 3
        bad = int((st != 0).sum())
        print(f"inflate_batch_dev ck={ckind} wall {dt*1e3:.2f} ms -> {U/dt/1e9:.2f} GB/s uncompressed; bad={bad}")
    out = ddst.cpu().numpy()
    okc = all(zlib.crc32(out[int(dst_off[i]):int(dst_off[i]) + len(datas[i])]) == zlib.crc32(datas[i]) for i in range(0, nmem, 37))
    print("inflate outputs ok:", okc, "crc ok:", all(int(ck[i]) == zlib.crc32(datas[i]) for i in range(nmem)))

if "deflate" in what:
    nmem = 2000
    sizes = synth.member_sizes(nmem, seed=7)
    datas = [synth.text_v1(1000 + i, int(s)) for i, s in enumerate(sizes)]
    U = sum(x.size for x in datas)
    for lvl in ("fast", "default", "best"):
        t0 = time.time()
        res = ctx.deflate_batch(datas, lvl, 2)
        dt = time.time() - t0
        Cb = sum(r[1].size for r in res)
        ok = all(zlib.decompress(r[1].tobytes(), -15) == d.tobytes() for r, d in list(zip(res, datas))[::53])
        print(f"deflate_batch {lvl}: e2e wall {dt*1e3:.1f} ms {U/dt/1e9:.2f} GB/s ratio {Cb/U:.4f} ok={ok}")
print("launches", ctx.launches)
