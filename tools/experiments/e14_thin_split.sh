set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/e14_pytest.txt 2>&1; tail -2 $O/e14_pytest.txt
python - <<'P'
import sys, time, zlib
sys.path.insert(0, '.')
from zipc_b200 import synth, _lib
from zipc_b200 import zipc_deflate as zd
import os
ctx = zd.Context(0)
for kib in (300, 512, 1024, 1900):
    x = synth.text_v1(9, kib << 10)
    for env in (None, str(1 << 40)):
        if env: os.environ["ZIPC_B200_SPLIT_MIN"] = env
        else: os.environ.pop("ZIPC_B200_SPLIT_MIN", None)
        ctx.deflate_batch([x], "default", _lib.CK_CRC32)
        t0 = time.perf_counter()
        for _ in range(3): st, cs, ck = ctx.deflate_batch([x], "default", _lib.CK_CRC32)[0]
        dt = (time.perf_counter() - t0) / 3
        assert st == 0 and ck == zlib.crc32(x) and zlib.decompress(bytes(cs), -15) == x.tobytes()
        print(f"{kib} KiB member alone, {'one CTA' if env else 'split    '}: {dt*1e3:.2f} ms per call, {len(cs)} bytes")
os.environ.pop("ZIPC_B200_SPLIT_MIN", None)
xs = [synth.text_v1(100 + i, 1 << 20) for i in range(50)]
for env in (None, str(1 << 40)):
    if env: os.environ["ZIPC_B200_SPLIT_MIN"] = env
    else: os.environ.pop("ZIPC_B200_SPLIT_MIN", None)
    ctx.deflate_batch(xs, "default", _lib.CK_CRC32)
    t0 = time.perf_counter(); res = ctx.deflate_batch(xs, "default", _lib.CK_CRC32); dt = time.perf_counter() - t0
    print(f"50 members of 1 MiB, {'one CTA each' if env else 'split'}: {dt*1e3:.2f} ms per call, {sum(len(r[1]) for r in res)} bytes")
P
