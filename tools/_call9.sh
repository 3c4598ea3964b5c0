#!/bin/bash
timeout 300 python -m pytest tests/test_eigen_gpu.py -m gpu -q -s 2>&1 | tail -12 > gpurun_out/c9_eigen.log
timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_dropin_models_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/c9_dropin_tests.log
timeout 300 python tools/dropin_bench.py 100000 m0 > gpurun_out/c9_dropin_m0.log 2>&1
timeout 600 python tools/dropin_bench.py 50000 sites > gpurun_out/c9_dropin_sites.log 2>&1
tail -12 gpurun_out/c9_eigen.log; tail -8 gpurun_out/c9_dropin_tests.log; cat gpurun_out/c9_dropin_m0.log gpurun_out/c9_dropin_sites.log
