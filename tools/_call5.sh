mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/c5_pytest.log 2>&1
tail -2 gpurun_out/c5_pytest.log
for v in 2 23 27 13 3 4 21; do PB200_NUC_PPT=$v timeout 300 python tools/quick_bench.py cfg5 1000000 2>&1 | tail -1; done
