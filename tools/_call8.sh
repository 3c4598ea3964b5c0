#!/bin/bash
timeout 300 python -m pytest tests/test_eigen_gpu.py -m gpu -q -s 2>&1 | tail -12 > gpurun_out/c8_eigen.log
timeout 300 python tools/dropin_bench.py 100000 > gpurun_out/c8_dropin.log 2>&1
tail -12 gpurun_out/c8_eigen.log; cat gpurun_out/c8_dropin.log
