"""Probe: how fast is ONE stream on one warp with the GPU otherwise idle, and how much of that is the copy phase (E)?
ZIPC_B200_LIB=.../libzipc_b200_noe.so (built with -DZB_PROBE_NO_E=1) skips the copy phase (output garbage, timing only)."""
import ctypes as C, os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if not os.environ.get("E6_PAR"):
    os.environ["ZIPC_B200_PAR_MIN"] = "0"   # no many-warp decoding: the one-warp decoder is what is measured (E6_PAR=1: the library decides)
import torch
from zipc_b200 import synth
from zipc_b200 import zipc_deflate as zd
ctx = zd.Context(0); L = ctx.L
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
for kib in (64, 256, 1024):
    for nstreams in (1, 16, 148):
        data = synth.text_v1(3, kib << 10)
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        zs = np.frombuffer(c.compress(data.tobytes()) + c.flush(), dtype=np.uint8)
        stride = (zs.size + 63) & ~63; ostride = (data.size + 63) & ~63
        host = np.zeros(stride * nstreams + 64, np.uint8)
        for i in range(nstreams): host[i * stride:i * stride + zs.size] = zs
        dsrc = torch.from_numpy(host).cuda()
        ddst = torch.empty(ostride * nstreams + 64, dtype=torch.uint8, device="cuda")
        coff = (np.arange(nstreams) * stride).astype(np.uint64); clen = np.full(nstreams, zs.size, np.uint64)
        ooff = (np.arange(nstreams) * ostride).astype(np.uint64); slen = np.full(nstreams, data.size, np.uint64)
        dl = np.zeros(nstreams, np.uint64); ck = np.zeros(nstreams, np.uint32); st = np.zeros(nstreams, np.int32)
        def fn():
            rc = L.zipc_b200_inflate_batch_dev(ctx.h, 0, 0, nstreams, dsrc.data_ptr(), P(coff, C.c_size_t), P(clen, C.c_size_t), ddst.data_ptr(),
                                               P(ooff, C.c_size_t), P(slen, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
            assert rc == 0, rc
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"{kib:5d} KiB x {nstreams:3d} streams ({zs.size >> 10} KiB compressed each): {dt*1e3:7.2f} ms per call = {data.size/dt/1e6:7.1f} MB/s per stream")
