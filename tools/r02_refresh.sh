#!/bin/bash
# Refresh of the evidence after the last changes of round 2 (lanes + spread speculative decode + Adler-32 through the many-warp
# decoder, split thresholds, stored kernel): GPU suite, bench line, launch list, fuzzers, corpus, sanitizers over the new paths.
# Outputs under gpurun_out/r02/.
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -2 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout 600 python tools/fuzz_inflate_large.py 300 13 > $O/fuzz_inflate_large.txt 2>&1; tail -2 $O/fuzz_inflate_large.txt
timeout 600 python tools/fuzz_deflate.py 2000 14 > $O/fuzz_deflate.txt 2>&1; tail -3 $O/fuzz_deflate.txt
timeout 300 python tools/corpus_wheels.py > $O/corpus_wheels.txt 2>&1; tail -1 $O/corpus_wheels.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_inflate_parallel.py tests/test_gpu_deflate.py -m gpu -x -q -k "mid_sized or adler32_take or zlib_compress_of_large or level_none or batch_mix" > $O/sanitizer_memcheck_new_paths.txt 2>&1; tail -3 $O/sanitizer_memcheck_new_paths.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python __graft_entry__.py smoke > $O/sanitizer_race_smoke.txt 2>&1; tail -2 $O/sanitizer_race_smoke.txt
timeout 600 compute-sanitizer --tool synccheck --num-cuda-barriers 16384 --print-limit 20 python __graft_entry__.py smoke > $O/sanitizer_sync_smoke.txt 2>&1; tail -2 $O/sanitizer_sync_smoke.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1; tail -1 $O/bench_under_ncu.log | cut -c1-120
