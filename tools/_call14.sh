#!/bin/bash
timeout 900 python -m pytest tests/test_dropin_models_gpu.py -m gpu -q -k "method1" 2>&1 | tail -40 > gpurun_out/c14_method1.log
timeout 900 python tools/method1_bench.py 32 2000 > gpurun_out/c14_method1_bench.log 2>&1
timeout 900 python tools/method1_bench.py 32 10000 > gpurun_out/c14_method1_bench10k.log 2>&1
tail -30 gpurun_out/c14_method1.log; cat gpurun_out/c14_method1_bench.log gpurun_out/c14_method1_bench10k.log
