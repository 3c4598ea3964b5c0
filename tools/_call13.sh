#!/bin/bash
timeout 600 python -m pytest tests/test_cond_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/c13_cond.log
timeout 900 python tools/method1_bench.py 32 2000 > gpurun_out/c13_method1_bench.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_multi_gpu.py --deselect tests/test_cond_gpu.py 2>&1 | tail -15 > gpurun_out/c13_pytest.log
tail -6 gpurun_out/c13_cond.log; cat gpurun_out/c13_method1_bench.log; tail -6 gpurun_out/c13_pytest.log
