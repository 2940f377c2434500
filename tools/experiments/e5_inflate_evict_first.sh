set -x
O=gpurun_out/r02b
mkdir -p $O
bash tools/ab_inflate.sh default ef default ef > $O/e5_ab.txt 2>&1; cat $O/e5_ab.txt
ZIPC_B200_LIB=$PWD/zipc_b200/libzipc_b200_ef.so timeout 300 python -m pytest tests/test_gpu_inflate.py -q -m gpu -x > $O/e5_pytest.txt 2>&1; tail -2 $O/e5_pytest.txt
for tag in default ef; do
  lib=""; [ "$tag" != default ] && lib=$PWD/zipc_b200/libzipc_b200_$tag.so
  ZIPC_B200_LIB=$lib timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:inflate_kernel -s 2 -c 1 python bench.py --workload inflate --no-also --steps 1 --warmup 2 > $O/e5_ncu_$tag.txt 2>&1; grep -E "dram__|lts__|gpu__time" $O/e5_ncu_$tag.txt
done
