set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_deflate.py -q -m gpu -x > $O/e4_pytest.txt 2>&1; tail -2 $O/e4_pytest.txt
for L in default fast; do ZIPC_B200_LIB=$PWD/zipc_b200/libzipc_b200_timing.so timeout 200 python tools/deflate_phases.py $L 3000 > $O/e4_phases_$L.txt 2>&1; head -1 $O/e4_phases_$L.txt; grep -E "shallow|wait_for" $O/e4_phases_$L.txt; done
timeout 300 python bench.py --workload deflate --no-also --steps 5 --warmup 3 > $O/e4_bench_deflate.json 2> $O/e4_bench_deflate.err; cut -c1-200 $O/e4_bench_deflate.json
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python __graft_entry__.py smoke > $O/e4_race_smoke.txt 2>&1; tail -4 $O/e4_race_smoke.txt
timeout 600 compute-sanitizer --tool synccheck --num-cuda-barriers 16384 --print-limit 20 python __graft_entry__.py smoke > $O/e4_sync_smoke.txt 2>&1; tail -4 $O/e4_sync_smoke.txt
