set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_inflate_parallel.py tests/test_gpu_inflate.py tests/test_gpu_deflate.py tests/test_gpu_checksums.py -q -m gpu -x > $O/e16_pytest.txt 2>&1; tail -15 $O/e16_pytest.txt
python - <<'P'
import sys, time, zlib
sys.path.insert(0, '.')
from zipc_b200 import synth, _lib
from zipc_b200 import zipc_deflate as zd
ctx = zd.Context(0); zd.set_default_context(ctx)
data = synth.text_v1(1, 64 << 20).tobytes()
z = zlib.compress(data, 6)
for mode, name in ((_lib.ADLER_REF_COMPAT, "REF_COMPAT"), (_lib.ADLER_RFC1950, "RFC1950")):
    zd.zlib_decompress(z, decompressed_size=len(data), adler_mode=mode)
    t0 = time.perf_counter(); r = zd.zlib_decompress(z, decompressed_size=len(data), adler_mode=mode); dt = time.perf_counter() - t0
    out, ad = r.get_ok()
    assert out == data and ad == zlib.adler32(data)
    print(f"zlib_decompress of one 64 MiB stream ({name}): {dt*1e3:.1f} ms = {len(data)/dt/1e9:.2f} GB/s (pageable buffers, python mirror); parallel/fallback {ctx.parallel_streams}")
P
