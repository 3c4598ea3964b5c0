#!/bin/bash
timeout 300 python -m pytest tests/test_eigen_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/c7_eigen.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_multi_gpu.py --deselect tests/test_eigen_gpu.py 2>&1 | tail -40 > gpurun_out/c7_pytest.log
timeout 300 python tools/dropin_bench.py 100000 > gpurun_out/c7_dropin.log 2>&1
tail -25 gpurun_out/c7_eigen.log; tail -15 gpurun_out/c7_pytest.log; cat gpurun_out/c7_dropin.log
