set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02/pytest_gpu.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deflate_kernel -s 1 -c 1 -f -o gpurun_out/r02/deflate_default python tools/prof_codec.py deflate 1200 default > gpurun_out/r02/ncu_deflate.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deflate_kernel -s 1 -c 1 -f -o gpurun_out/r02/deflate_fast python tools/prof_codec.py deflate 1200 fast > gpurun_out/r02/ncu_deflate_fast.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -s 1 -c 1 -f -o gpurun_out/r02/inflate python tools/prof_codec.py inflate 1200 default > gpurun_out/r02/ncu_inflate.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02/bench_under_ncu.log 2>&1
ls -la gpurun_out/r02
