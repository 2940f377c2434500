"""Scratch: run one codec batch on device-resident data (for ncu). usage: prof_codec.py deflate|inflate [n] [level]"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd
which = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 1200; level = sys.argv[3] if len(sys.argv) > 3 else "default"
ctx = zd.Context(0); L = ctx.L
sizes = synth.member_sizes(n, seed=3)
datas = [synth.text_v1(1000 + i, int(s)) for i, s in enumerate(sizes)]
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
def pack(items):
    offs = np.zeros(len(items), dtype=np.uint64); t = 0
    for i, x in enumerate(items):
        offs[i] = t; t += (x.size + 15) & ~15
    host = np.zeros(t + 64, dtype=np.uint8)
    for i, x in enumerate(items): host[int(offs[i]):int(offs[i]) + x.size] = x
    return torch.from_numpy(host).cuda(), offs, np.array([x.size for x in items], dtype=np.uint64)
dsrc, soff, slen = pack(datas)
dl = np.zeros(n, dtype=np.uint64); ck = np.zeros(n, dtype=np.uint32); st = np.zeros(n, dtype=np.int32)
lvl = {"fast": 1, "default": 2, "best": 3}[level]
if which == "deflate":
    cap = np.array([(L.zipc_b200_deflate_bound(int(x)) + 15) & ~15 for x in slen], dtype=np.uint64)
    doff = np.concatenate([[0], np.cumsum(cap)[:-1]]).astype(np.uint64)
    ddst = torch.empty(int(cap.sum()) + 64, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        rc = L.zipc_b200_deflate_batch_dev(ctx.h, lvl, 0, 0, n, dsrc.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t), ddst.data_ptr(), P(doff, C.c_size_t), P(cap, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
        assert rc == 0 and (st == 0).all()
    print("ratio", dl.sum() / slen.sum())
else:
    res = ctx.deflate_batch(datas, level, 0)
    streams = [r[1] for r in res]
    dcs, coff, clen = pack(streams)
    ddst = torch.empty(int(soff[-1] + slen[-1]) + 64, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        rc = L.zipc_b200_inflate_batch_dev(ctx.h, 0, 0, n, dcs.data_ptr(), P(coff, C.c_size_t), P(clen, C.c_size_t), ddst.data_ptr(), P(soff, C.c_size_t), P(slen, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
        assert rc == 0 and (st == 0).all()
print("done")
