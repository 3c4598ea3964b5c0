set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
for v in 2 4 3 1; do PB200_NUC_PPT=$v timeout 300 python tools/quick_bench.py cfg5 1000000 >> gpurun_out/c1_cfg5.log 2>&1; done
timeout 300 python tools/quick_bench.py cfg4 >> gpurun_out/c1_cfg4.log 2>&1
timeout 300 python tools/quick_bench.py cfg3 >> gpurun_out/c1_cfg3.log 2>&1
timeout 300 python tools/quick_bench.py cfg2 >> gpurun_out/c1_cfg2.log 2>&1
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_cfg5.log gpurun_out/c1_cfg4.log gpurun_out/c1_cfg3.log gpurun_out/c1_cfg2.log
