#!/bin/bash
# Short refresh of the evidence after the last inflate changes of round 2 (lanes, spread speculative decode): GPU suite, bench line,
# launch list, the large-stream fuzzer, the corpus, memcheck over the inflate tests.  Outputs under gpurun_out/r02/.
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout 600 python tools/fuzz_inflate_large.py 300 12 > $O/fuzz_inflate_large.txt 2>&1; tail -2 $O/fuzz_inflate_large.txt
timeout 300 python tools/corpus_wheels.py > $O/corpus_wheels.txt 2>&1; tail -1 $O/corpus_wheels.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_inflate_parallel.py tests/test_gpu_inflate.py -m gpu -x -q -k "not 64mib and not large_zlib" > $O/sanitizer_memcheck_inflate.txt 2>&1; tail -3 $O/sanitizer_memcheck_inflate.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1; tail -2 $O/bench_under_ncu.log | cut -c1-200
