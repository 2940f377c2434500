set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/e11_pytest.txt 2>&1; tail -3 $O/e11_pytest.txt
E6_PAR=1 python tools/experiments/e6_lone_stream_probe.py > $O/e11_probe_par.txt 2>&1; cat $O/e11_probe_par.txt
python tools/experiments/e9_few_large.py 2>&1 | head -1 | cut -c1-80,290-420
timeout 300 python tools/corpus_wheels.py > $O/e11_corpus.txt 2>&1; tail -1 $O/e11_corpus.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_inflate_parallel.py -m gpu -x -q -k "mid_sized or batch_mix" > $O/e11_memcheck_lanes.txt 2>&1; tail -3 $O/e11_memcheck_lanes.txt
