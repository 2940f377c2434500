set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_inflate_parallel.py tests/test_gpu_zip.py tests/test_gpu_tool.py -q -m gpu -x > $O/e7_pytest.txt 2>&1; tail -3 $O/e7_pytest.txt
E6_PAR=1 python tools/experiments/e6_lone_stream_probe.py > $O/e7_probe_par.txt 2>&1; cat $O/e7_probe_par.txt
timeout 300 python tools/corpus_wheels.py > $O/e7_corpus.txt 2>&1; tail -1 $O/e7_corpus.txt
ZIPC_B200_PAR_MIN=262144 timeout 300 python tools/corpus_wheels.py > $O/e7_corpus_old.txt 2>&1; tail -1 $O/e7_corpus_old.txt
timeout 300 python tools/fuzz_inflate_large.py 100 9 > $O/e7_fuzz_large.txt 2>&1; tail -2 $O/e7_fuzz_large.txt
