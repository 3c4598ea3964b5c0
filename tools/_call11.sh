#!/bin/bash
timeout 600 python -m pytest tests/test_cond_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/c11_cond.log
tail -40 gpurun_out/c11_cond.log
