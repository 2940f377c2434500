#!/bin/bash
# Final evidence pass of round 2 (run on the GPU box through gpurun): tools/r02_profile_pass.sh, then the fuzzers and the
# sanitizers on the final kernels.  Outputs under gpurun_out/r02/.
O=gpurun_out/r02
mkdir -p $O
bash tools/r02_profile_pass.sh > $O/profile_pass.log 2>&1
tail -3 $O/pytest_gpu.txt
timeout 600 python tools/fuzz_deflate.py 3000 11 > $O/fuzz_deflate.txt 2>&1; tail -3 $O/fuzz_deflate.txt
timeout 600 python tools/fuzz_inflate.py 60000 778 > $O/fuzz_inflate.txt 2>&1; tail -3 $O/fuzz_inflate.txt
timeout 600 python tools/fuzz_inflate_large.py 200 6 > $O/fuzz_inflate_large.txt 2>&1; tail -2 $O/fuzz_inflate_large.txt
timeout 300 python tools/corpus_wheels.py > $O/corpus_wheels.txt 2>&1; tail -2 $O/corpus_wheels.txt
bash tools/r02_sanitizer_pass.sh > $O/sanitizer_pass.log 2>&1
tail -4 $O/sanitizer_memcheck.txt; tail -3 $O/sanitizer_race_adler.txt; tail -3 $O/sanitizer_race_smoke.txt; tail -3 $O/sanitizer_sync_smoke.txt
