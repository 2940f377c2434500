"""Fuzz of the many-warp decoder of ONE large stream (block-start search, speculative decode, chain check, re-cuts) against the
oracle: mutated streams of 1-3 MiB (zlib at several levels, this library's own split members, streams with stored / fixed
blocks in the middle), with and without a size limit; compares status, bytes and CRC.
python tools/fuzz_inflate_large.py [count] [seed]"""
import os, random, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import zipc_oracle as zo
from zipc_b200 import _lib, synth
from zipc_b200 import zipc_deflate as zd

count = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = zd.Context(0)
text = synth.text_v1(9, 3 << 20).tobytes()
noise = synth.rand_v1(10, 300_000).tobytes()

def z(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()

mixed = zlib.compressobj(6, zlib.DEFLATED, -15)
mixed_s = mixed.compress(text[:700_000]) + mixed.flush(zlib.Z_FULL_FLUSH) + mixed.compress(noise) + mixed.flush(zlib.Z_SYNC_FLUSH) + mixed.compress(text[700_000:2_000_000]) + mixed.flush()
bases = [z(text, 6), z(text[:(1 << 20) + 5], 1), z(text[: 2 << 20], 9), mixed_s,
         bytes(ctx.deflate_batch([text], "default", 0)[0][1]),          # split into primed segments (>= 2 MiB)
         bytes(ctx.deflate_batch([text[:(2 << 20) + 77]], "fast", 0)[0][1])]
assert all(len(b) >= 262144 for b in bases), [len(b) for b in bases]

def mutate(b):
    b = bytearray(b)
    k = rnd.randrange(6)
    if k == 0: del b[rnd.randrange(len(b) // 2, len(b)):]
    elif k == 1:
        for _ in range(rnd.randrange(1, 4)): b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
    elif k == 2:
        p = rnd.randrange(len(b)); b[p:p + rnd.randrange(1, 9)] = rnd.randbytes(rnd.randrange(1, 9))
    elif k == 3:
        p = rnd.randrange(len(b)); b[p:p] = rnd.randbytes(rnd.randrange(1, 5))
    elif k == 4:
        p = rnd.randrange(len(b) - 70000); del b[p:p + rnd.randrange(1, 70000)]
    return bytes(b)   # k == 5: unchanged

bad = 0
par0, fb0 = ctx.parallel_streams
stats = {}
for it in range(count):
    s = mutate(rnd.choice(bases))
    limit = rnd.choice([None, None, 0, 1 << 20, (2 << 20) + 77, 3 << 20, (3 << 20) - 1])
    st, out, ck = ctx.inflate_batch([s], [limit], _lib.CK_CRC32)[0]
    try:
        eo, ec = zo.inflate_and_crc(s, limit, zo.CRC_CRC32); est = 0
    except zo.OracleError as e:
        est, eo, ec = e.status, b"", 0
    stats[est] = stats.get(est, 0) + 1
    if st != est or (st == 0 and (out.tobytes() != eo or ck != ec)):
        bad += 1
        if bad <= 5: print("MISMATCH", it, "gpu", st, "oracle", est, "len", len(s), "limit", limit)
par1, fb1 = ctx.parallel_streams
print("oracle status histogram", stats)
print("streams", count, "mismatches", bad, "| decoded in parallel", par1 - par0, "| handed to the one-warp decoder", fb1 - fb0)
sys.exit(1 if bad else 0)
