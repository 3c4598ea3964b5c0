#!/bin/bash
timeout 600 python -m pytest tests/test_cond_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/c12_cond.log
timeout 900 python -m pytest tests/test_dropin_models_gpu.py -m gpu -q -k "method1" 2>&1 | tail -60 > gpurun_out/c12_method1.log
tail -25 gpurun_out/c12_cond.log; tail -45 gpurun_out/c12_method1.log
