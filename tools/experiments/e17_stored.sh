set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_zip.py -q -m gpu -x > $O/e17_pytest.txt 2>&1; tail -3 $O/e17_pytest.txt
python - <<'P'
import sys, time, zlib, json
sys.path.insert(0, '.')
import bench
from zipc_b200 import synth, _lib
h = bench.Harness(0)
datas = bench.make_members(synth, 10000, 1000)
r = bench.run_deflate(h, datas, "none", 3, 3, e2e=False)
print("C4 members at level None: %.1f GB/s of input device-resident (%.3f ms per batch), kernel_ms %s" % (r["units"] * r["steps"] / (r["total_ms"] / 1e3) / 1e9, r["total_ms"] / r["steps"], r["kernel_ms"]))
x = synth.text_v1(5, 64 << 20)
h.ctx.deflate_batch([x], "none", _lib.CK_CRC32)
t0 = time.perf_counter(); st, cs, ck = h.ctx.deflate_batch([x], "none", _lib.CK_CRC32)[0]; dt = time.perf_counter() - t0
assert st == 0 and ck == zlib.crc32(x) and zlib.decompress(bytes(cs), -15) == x.tobytes()
print("one 64 MiB member at level None through the mirror (pageable): %.1f ms" % (dt * 1e3))
P
