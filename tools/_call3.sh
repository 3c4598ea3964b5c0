mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.log 2>&1
tail -3 gpurun_out/c3_pytest.log
PB200_NUC_KERNEL=v1 timeout 300 python tools/quick_bench.py cfg5 1000000 2>&1 | tail -1
for v in 2 4 41 21 3 1; do PB200_NUC_PPT=$v timeout 300 python tools/quick_bench.py cfg5 1000000 2>&1 | tail -1; done
