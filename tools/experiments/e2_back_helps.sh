set -x
mkdir -p gpurun_out/r02b
timeout 300 python -m pytest tests/test_gpu_deflate.py -q -m gpu -x > gpurun_out/r02b/e2_pytest.txt 2>&1; tail -2 gpurun_out/r02b/e2_pytest.txt
for H in 0 1; do
  export ZIPC_B200_LIB=$PWD/zipc_b200/libzipc_b200_h$H.so
  for L in default fast; do timeout 200 python tools/deflate_phases.py $L 3000 > gpurun_out/r02b/e2_phases_h${H}_$L.txt 2>&1; head -1 gpurun_out/r02b/e2_phases_h${H}_$L.txt; grep -E "shallow|wait_for" gpurun_out/r02b/e2_phases_h${H}_$L.txt; done
done
unset ZIPC_B200_LIB
timeout 300 python bench.py --workload deflate --no-also --steps 5 --warmup 3 > gpurun_out/r02b/e2_bench_deflate.json 2> gpurun_out/r02b/e2_bench_deflate.err; cut -c1-400 gpurun_out/r02b/e2_bench_deflate.json
