"""Scratch: time the C3 inflate batch under different launch knobs (env read at each launch)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ctx = zd.Context(0); L = ctx.L
sizes = synth.member_sizes(n, seed=3)
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex:
    datas = list(ex.map(lambda a: synth.text_v1(1000 + a[0], int(a[1])), enumerate(sizes)))
res = ctx.deflate_batch(datas, "default", 2)
streams = [r[1] for r in res]; crcs = np.array([r[2] for r in res], dtype=np.uint32)
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
def pack(items):
    offs = np.zeros(len(items), dtype=np.uint64); t = 0
    for i, x in enumerate(items):
        offs[i] = t; t += (x.size + 15) & ~15
    host = np.zeros(t + 64, dtype=np.uint8)
    for i, x in enumerate(items): host[int(offs[i]):int(offs[i]) + x.size] = x
    return torch.from_numpy(host).cuda(), offs, np.array([x.size for x in items], dtype=np.uint64)
dcs, coff, clen = pack(streams)
_, soff, slen = pack(datas)
U = int(slen.sum())
ddst = torch.empty(int(soff[-1] + slen[-1]) + 64, dtype=torch.uint8, device="cuda")
dl = np.zeros(n, dtype=np.uint64); ck = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)
def run():
    rc = L.zipc_b200_inflate_batch_dev(ctx.h, 2, 0, n, dcs.data_ptr(), P(coff, C.c_size_t), P(clen, C.c_size_t), ddst.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
    assert rc == 0 and (st == 0).all() and (ck == crcs).all()
L.zipc_b200_ctx_profile(ctx.h, 1)
for key, val in [("", ""), ("ZIPC_B200_INFLATE_OVERSUB", "2"), ("ZIPC_B200_INFLATE_OVERSUB", "3"), ("ZIPC_B200_INFLATE_OVERSUB", "4"),
                 ("ZIPC_B200_INFLATE_WARPS", "16"), ("ZIPC_B200_INFLATE_WARPS", "12"), ("ZIPC_B200_INFLATE_WARPS", "6"), ("ZIPC_B200_INFLATE_WARPS", "4"), ("ZIPC_B200_INFLATE_WARPS", "3")]:
    for k in ("ZIPC_B200_INFLATE_OVERSUB", "ZIPC_B200_INFLATE_WARPS"): os.environ.pop(k, None)
    if key: os.environ[key] = val
    run(); ts = []
    for _ in range(3):
        run(); ts.append(L.zipc_b200_ctx_kernel_ms(ctx.h))
    print(f"{key or 'default':28s} {val:3s} inflate kernel {min(ts):8.3f} ms -> {U/min(ts)/1e6:7.2f} GB/s", flush=True)
