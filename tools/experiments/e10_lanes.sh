set -x
O=gpurun_out/r02b
mkdir -p $O
python tools/experiments/e9_few_large.py > $O/e10_after.txt 2> $O/e10_after.err; cat $O/e10_after.txt; tail -3 $O/e10_after.err
ZIPC_B200_PAR_LANES=1 python tools/experiments/e9_few_large.py > $O/e10_lanes1.txt 2>&1; head -c 600 $O/e10_lanes1.txt
for Ln in 4 8 16; do ZIPC_B200_PAR_LANES=$Ln python tools/experiments/e9_few_large.py 2>&1 | head -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('lanes $Ln', d['inflate_e2e_GBps'], d['deflate_e2e_GBps'])"; done
timeout 600 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_inflate_parallel.py tests/test_gpu_zip.py tests/test_gpu_tool.py tests/test_gpu_deflate.py -q -m gpu -x > $O/e10_pytest.txt 2>&1; tail -3 $O/e10_pytest.txt
E6_PAR=1 python tools/experiments/e6_lone_stream_probe.py > $O/e10_probe_par.txt 2>&1; cat $O/e10_probe_par.txt
timeout 300 python tools/fuzz_inflate_large.py 150 10 > $O/e10_fuzz_large.txt 2>&1; tail -2 $O/e10_fuzz_large.txt
