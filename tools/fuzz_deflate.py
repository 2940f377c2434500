"""One-off fuzz of the batch deflate: inputs stitched from text, runs, random bytes, repeats at assorted distances (incl.
32768 and beyond the window), sizes 0 .. 400 KB, every level; each output must be a valid stream that zlib inflates back to
the input, with the right CRC-32, and stay within deflate_bound.  python tools/fuzz_deflate.py [count] [seed]"""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

count = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = zd.Context(0)
text = synth.text_v1(9, 1 << 20)

def piece():
    k = int(rng.integers(0, 7)); n = int(rng.integers(1, 60000))
    if k == 0: o = int(rng.integers(0, text.size - n)); return text[o:o + n]
    if k == 1: return np.full(n, int(rng.integers(0, 256)), np.uint8)
    if k == 2: return rng.integers(0, 256, n, dtype=np.uint8)
    if k == 3: return rng.integers(0, int(rng.integers(2, 6)), n, dtype=np.uint8)
    if k == 4: p = rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8); return np.tile(p, n // p.size + 1)[:n]
    if k == 5: return np.minimum(rng.exponential(float(rng.uniform(0.3, 3)), n), 255).astype(np.uint8)
    return np.arange(n, dtype=np.uint32).astype(np.uint8)

def make():
    if rng.random() < 0.03: return np.zeros(int(rng.integers(0, 5)), np.uint8)
    parts = [piece() for _ in range(int(rng.integers(1, 8)))]
    if rng.random() < 0.3:   # a far repeat: the same block again 32768 +- a few bytes (or further) later
        gap = rng.integers(0, 256, int(rng.choice([32768 - 300, 32767, 32768, 32769, 40000])) - min(parts[0].size, 300), dtype=np.uint8)
        parts = [parts[0][:300], gap, parts[0][:300]] + parts[1:]
    return np.concatenate(parts)[:400000]

datas = [make() for _ in range(count)]
bad = 0
for level in ("fast", "default", "best", "none"):
    res = ctx.deflate_batch(datas, level, _lib.CK_CRC32)
    tot_c = 0
    for i, (st, out, ck) in enumerate(res):
        d = datas[i].tobytes()
        ok = st == 0 and ck == zlib.crc32(d) and out.size <= ctx.L.zipc_b200_deflate_bound(len(d))
        if ok:
            try: ok = zlib.decompress(out.tobytes(), -15) == d
            except zlib.error: ok = False
        if not ok:
            bad += 1
            if bad <= 5: print("FAIL", level, i, "status", st, "len", len(d))
        tot_c += out.size
    print("level %-8s ratio %.4f" % (level, tot_c / max(1, sum(d.size for d in datas))), flush=True)
print("inputs", count, "bytes", sum(d.size for d in datas), "failures", bad)
sys.exit(1 if bad else 0)
