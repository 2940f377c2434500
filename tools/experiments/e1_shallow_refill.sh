set -x
mkdir -p gpurun_out/r02b
for T in 0 8 16 24; do
  export ZIPC_B200_LIB=$PWD/zipc_b200/libzipc_b200_t$T.so
  if [ $T != 0 ]; then timeout 300 python -m pytest tests/test_gpu_deflate.py -q -m gpu -x -k "model_equality or primed or ragged" > gpurun_out/r02b/e1_pytest_t$T.txt 2>&1; tail -2 gpurun_out/r02b/e1_pytest_t$T.txt; fi
  for L in default fast; do timeout 200 python tools/deflate_phases.py $L 3000 > gpurun_out/r02b/e1_phases_t${T}_$L.txt 2>&1; head -1 gpurun_out/r02b/e1_phases_t${T}_$L.txt; grep shallow gpurun_out/r02b/e1_phases_t${T}_$L.txt; done
done
