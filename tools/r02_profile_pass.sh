#!/bin/bash
# Round-2 evidence pass (run on the GPU box through gpurun): tests, both bench arms, launch list, ncu captures of the
# dominant kernels ON THE BENCH COMMAND.  Outputs under gpurun_out/r02/; the summaries are made from them with
# tools/ncu_summary.py and kept under profiles/.
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
for w in deflate inflate crc32; do
  case $w in deflate) k=deflate_kernel;; inflate) k=inflate_kernel;; crc32) k=crc32_tiles_kernel;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/bench_$w python bench.py --workload $w --no-also --steps 1 --warmup 2 > $O/ncu_bench_$w.log 2>&1
done
ls -la $O
