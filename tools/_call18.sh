#!/bin/bash
timeout 600 python -m pytest tests/test_cond_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/c19_cond.log
timeout 900 python -m pytest tests/test_dropin_models_gpu.py -m gpu -q -k "method1" 2>&1 | tail -12 > gpurun_out/c19_method1.log
timeout 300 python tools/cond_profile.py 50000 4 > gpurun_out/c19_cond_profile.log 2>&1
timeout 300 python tools/method1_bench.py 32 10000 > gpurun_out/c19_method1_bench.log 2>&1
tail -5 gpurun_out/c19_cond.log; tail -5 gpurun_out/c19_method1.log; cat gpurun_out/c19_cond_profile.log; tail -1 gpurun_out/c19_method1_bench.log
