set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_zip.py -q -m gpu -x > $O/e13_pytest.txt 2>&1; tail -2 $O/e13_pytest.txt
python tools/experiments/e9_few_large.py 2>&1 | head -1 | cut -c1-60,290-420
python - <<'P'
import sys, time, zlib, json
sys.path.insert(0, '.')
import bench
h = bench.Harness(0)
d = bench.run_stream_c1(h)
print(json.dumps(d['one_member_batch_call']))
from zipc_b200 import synth
from zipc_b200 import _lib
for mib in (3, 5, 16):
    x = synth.text_v1(9, mib << 20)
    h.ctx.deflate_batch([x], "default", _lib.CK_CRC32)
    t0 = time.perf_counter()
    for _ in range(3): st, cs, ck = h.ctx.deflate_batch([x], "default", _lib.CK_CRC32)[0]
    dt = (time.perf_counter() - t0) / 3
    assert st == 0 and ck == zlib.crc32(x) and zlib.decompress(bytes(cs), -15) == x.tobytes()
    print(f"{mib} MiB member alone: {dt*1e3:.2f} ms per call (pageable buffers), ratio {len(cs)/x.size:.4f}")
P
