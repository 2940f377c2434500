set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_deflate.py tests/test_gpu_zip.py tests/test_gpu_tool.py tests/test_gpu_checksums.py -q -m gpu -x > $O/e15_pytest.txt 2>&1; tail -12 $O/e15_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -1
