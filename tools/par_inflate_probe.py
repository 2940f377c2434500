"""Tuning aid: one large foreign stream (zlib level L, no index) through zipc_b200_inflate_batch_dev / _batch."""
import ctypes as C, os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zipc_b200 import synth
from zipc_b200 import zipc_deflate as zd

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ctx = zd.Context(0)
L = ctx.L
data = synth.text_v1(1, mib << 20)
if level < 0:   # our own encoder: a large member, split into primed segments
    zs = np.array(ctx.deflate_batch([data], "default", 0)[0][1], dtype=np.uint8)
else:
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    zs = np.frombuffer(c.compress(data.tobytes()) + c.flush(), dtype=np.uint8)
P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
dsrc = torch.from_numpy(zs.copy()).cuda()
ddst = torch.empty(data.size + 64, dtype=torch.uint8, device="cuda")
off0 = np.zeros(1, dtype=np.uint64); clen = np.array([zs.size], dtype=np.uint64); slen = np.array([data.size], dtype=np.uint64)
dl = np.zeros(1, dtype=np.uint64); ck = np.zeros(1, dtype=np.uint32); st = np.zeros(1, dtype=np.int32)
def fn():
    rc = L.zipc_b200_inflate_batch_dev(ctx.h, 2, 0, 1, dsrc.data_ptr(), P(off0, C.c_size_t), P(clen, C.c_size_t), ddst.data_ptr(), P(off0, C.c_size_t),
                                       P(slen, C.c_size_t), P(dl, C.c_size_t), P(ck, C.c_uint32), P(st, C.c_int))
    assert rc == 0 and st[0] == 0, (rc, st[0])
for chunk in (sys.argv[3:] or ["8192"]):
    os.environ["ZIPC_B200_PAR_CHUNK"] = chunk
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    assert ck[0] == zlib.crc32(data) and int(dl[0]) == data.size
    print(f"{mib} MiB zlib -{level} ({zs.size/1e6:.1f} MB compressed), chunk {chunk}: device-resident {data.size/dt/1e9:.2f} GB/s ({dt*1e3:.1f} ms), parallel/fallback = {ctx.parallel_streams}")
os.environ["ZIPC_B200_PAR_MIN"] = "0"
t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"serial one-warp decode: {data.size/dt/1e9:.3f} GB/s ({dt*1e3:.0f} ms)")
