"""One-off fuzz of the batch inflate against the oracle: mutated streams (truncation, bit flips, spliced bytes, damaged
headers), with and without a size limit; compares status, bytes and CRC.  python tools/fuzz_inflate.py [count] [seed]"""
import os, random, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zipc_oracle as zo
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

count = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = zd.Context(0)
text = synth.text_v1(4, 60000).tobytes()
rng = np.random.default_rng(3)
skew = np.minimum(rng.exponential(0.7, 40000), 255).astype(np.uint8).tobytes()
bases = [zo.deflate(text, "default"), zo.deflate(text[:300], "fast"), zlib.compress(text, 1)[2:-4], zlib.compress(skew, 9)[2:-4],
         zo.deflate(rnd.randbytes(3000), "default"), zo.deflate(b"hellohello", "default"), zo.deflate(bytes(5000), "best")]
for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 9, strat)
    bases.append(c.compress(text[:8000]) + c.flush())
c = zlib.compressobj(6, zlib.DEFLATED, -15)
bases.append(b"".join(c.compress(text[i:i + 3000]) + c.flush(zlib.Z_SYNC_FLUSH) for i in range(0, 30000, 3000)) + c.flush())

def mutate(b):
    b = bytearray(b)
    k = rnd.randrange(5)
    if k == 0 and len(b) > 1: del b[rnd.randrange(1, len(b)):]
    elif k == 1:
        for _ in range(rnd.randrange(1, 4)): b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
    elif k == 2:
        p = rnd.randrange(len(b)); b[p:p + rnd.randrange(1, 9)] = rnd.randbytes(rnd.randrange(1, 9))
    elif k == 3: b[:rnd.randrange(1, 6)] = rnd.randbytes(rnd.randrange(1, 6))
    else:
        p = rnd.randrange(len(b)); b[p:p] = rnd.randbytes(rnd.randrange(1, 5))
    return bytes(b)

streams = [mutate(rnd.choice(bases)) for _ in range(count)]
bad = 0
for limits in (None, [rnd.choice([0, 1, 299, 300, 7999, 8000, 59999, 60000, 70000]) for _ in streams]):
    got = ctx.inflate_batch(streams, limits, _lib.CK_CRC32)
    stats = {}
    for i, (st, out, ck) in enumerate(got):
        try:
            eo, ec = zo.inflate_and_crc(streams[i], limits[i] if limits else None, zo.CRC_CRC32); est = 0
        except zo.OracleError as e:
            est, eo, ec = e.status, b"", 0
        stats[est] = stats.get(est, 0) + 1
        if st != est or (st == 0 and (out.tobytes() != eo or ck != ec)):
            bad += 1
            if bad <= 5: print("MISMATCH", i, "gpu", st, "oracle", est, "len", len(streams[i]), "limit", limits[i] if limits else None)
    print("limits" if limits else "no limits", "oracle status histogram", stats, flush=True)
print("streams", count, "mismatches", bad)
sys.exit(1 if bad else 0)
