#!/bin/bash
# compute-sanitizer evidence of round 2 (run through gpurun); outputs under gpurun_out/r02/
O=gpurun_out/r02
mkdir -p $O
SEL='not large and not progressive and not pipelined and not box_wide and not 64mib and not fuzz'
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/sanitizer_memcheck.txt 2>&1; tail -8 $O/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_checksums.py -m gpu -x -q -k "adler" > $O/sanitizer_race_adler.txt 2>&1; tail -6 $O/sanitizer_race_adler.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python __graft_entry__.py smoke > $O/sanitizer_race_smoke.txt 2>&1; tail -12 $O/sanitizer_race_smoke.txt
timeout 600 compute-sanitizer --tool synccheck --num-cuda-barriers 16384 --print-limit 20 python __graft_entry__.py smoke > $O/sanitizer_sync_smoke.txt 2>&1; tail -6 $O/sanitizer_sync_smoke.txt
