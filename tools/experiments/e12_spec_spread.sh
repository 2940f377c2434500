set -x
O=gpurun_out/r02b
mkdir -p $O
ZIPC_B200_PAR_DEBUG=1 python tools/par_inflate_probe.py 64 6 > $O/e12_c1.txt 2> $O/e12_c1.err; cat $O/e12_c1.txt; grep "pass 0" $O/e12_c1.err | tail -1
python tools/experiments/e9_few_large.py 2>&1 | head -1 | cut -c1-60,290-400
E6_PAR=1 python tools/experiments/e6_lone_stream_probe.py 2>&1 | grep -E " 1 streams| 16 streams"
timeout 600 python -m pytest tests/test_gpu_inflate_parallel.py tests/test_gpu_zip.py -q -m gpu -x > $O/e12_pytest.txt 2>&1; tail -2 $O/e12_pytest.txt
