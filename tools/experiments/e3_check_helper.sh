set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_zip.py -m gpu -x -q -k "zip64" > $O/e3_zip64_gpu.txt 2>&1; tail -3 $O/e3_zip64_gpu.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python __graft_entry__.py smoke > $O/e3_race_smoke.txt 2>&1; tail -6 $O/e3_race_smoke.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke > $O/e3_sync_smoke.txt 2>&1; tail -4 $O/e3_sync_smoke.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_deflate.py -m gpu -x -q -k "not large and not primed and not segmented" > $O/e3_memcheck_deflate.txt 2>&1; tail -4 $O/e3_memcheck_deflate.txt
