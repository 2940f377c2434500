set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_inflate_parallel.py tests/test_gpu_zip.py tests/test_gpu_tool.py tests/test_gpu_deflate.py -q -m gpu -x > $O/e8_pytest.txt 2>&1; tail -3 $O/e8_pytest.txt
E6_PAR=1 python tools/experiments/e6_lone_stream_probe.py > $O/e8_probe_par.txt 2>&1; cat $O/e8_probe_par.txt
timeout 300 python tools/corpus_wheels.py > $O/e8_corpus.txt 2>&1; tail -1 $O/e8_corpus.txt
timeout 300 python tools/fuzz_inflate_large.py 100 9 > $O/e8_fuzz_large.txt 2>&1; tail -2 $O/e8_fuzz_large.txt
