"""Throughput of the batch codecs on other kinds of data than text-v1 (device-resident, kernel-level):
python tools/data_kinds.py   -> one line per kind.  Looks for pathological cases (runs, incompressible data)."""
import ctypes as C, sys, time, zlib
import numpy as np, torch
sys.path.insert(0, ".")
from zipc_b200 import synth
from zipc_b200 import zipc_deflate as zd

ctx = zd.Context(0); L = ctx.L
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
rng = np.random.default_rng(5)
N, SZ = 2000, 256 << 10

def kinds():
    yield "text-v1", [synth.text_v1(100 + i, SZ) for i in range(N)]
    yield "zeros", [np.zeros(SZ, np.uint8) for _ in range(N)]
    yield "random", [synth.rand_v1(7 + i, SZ) for i in range(N)]
    yield "runs (RLE-like)", [np.repeat(rng.integers(0, 256, SZ // 64, dtype=np.uint8), 64) for _ in range(N)]
    yield "low-entropy bytes (no repeats)", [rng.integers(0, 4, SZ, dtype=np.uint8) for _ in range(N)]
    src = open("tests/golden/zip-docs.zip", "rb").read()
    import io, zipfile
    z = zipfile.ZipFile(io.BytesIO(src))
    docs = b"".join(z.read(n) for n in z.namelist())
    yield "zip-docs fixture (html/css/txt)", [np.frombuffer((docs * (SZ // len(docs) + 1))[i * 37:i * 37 + SZ], np.uint8) for i in range(N)]

def pack(items):
    offs = np.zeros(len(items), dtype=np.uint64); t = 0
    for i, x in enumerate(items):
        offs[i] = t; t += (x.size + 15) & ~15
    host = np.zeros(t + 64, dtype=np.uint8)
    for i, x in enumerate(items): host[int(offs[i]):int(offs[i]) + x.size] = x
    return torch.from_numpy(host).cuda(), offs, np.array([x.size for x in items], dtype=np.uint64)

for name, datas in kinds():
    U = sum(d.size for d in datas)
    res = ctx.deflate_batch(datas, "default", 2)
    assert all(r[0] == 0 for r in res)
    streams = [r[1] for r in res]; crcs = np.array([r[2] for r in res], dtype=np.uint32)
    Cb = sum(s.size for s in streams)
    assert zlib.decompress(streams[3].tobytes(), -15) == datas[3].tobytes()
    dsrc, soff, slen = pack(datas); dcs, coff, clen = pack(streams)
    n = len(datas)
    cap = np.array([L.zipc_b200_deflate_bound(int(x)) + 15 & ~15 for x in slen], dtype=np.uint64)
    doff = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint64)
    dd = torch.empty(int(cap.sum()) + 64, dtype=torch.uint8, device="cuda")
    di = torch.empty(int(soff[-1] + slen[-1]) + 64, dtype=torch.uint8, device="cuda")
    dl = np.zeros(n, dtype=np.uint64); ck = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)
    def defl():
        assert L.zipc_b200_deflate_batch_dev(ctx.h, 2, 2, 0, n, dsrc.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t), dd.data_ptr(),
                                             P(doff, C.c_size_t), P(cap, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int)) == 0
    def infl():
        assert L.zipc_b200_inflate_batch_dev(ctx.h, 2, 0, n, dcs.data_ptr(), P(coff, C.c_size_t), P(clen, C.c_size_t), di.data_ptr(),
                                             P(soff, C.c_size_t), P(slen, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int)) == 0
    out = []
    for fn in (defl, infl):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); fn(); torch.cuda.synchronize()
        out.append(U * 2 / (time.perf_counter() - t0) / 1e9)
        assert (st == 0).all() and (ck == crcs).all()
    print(f"{name:34s} ratio {Cb / U:7.4f}   deflate {out[0]:7.2f} GB/s   inflate {out[1]:7.2f} GB/s", flush=True)
