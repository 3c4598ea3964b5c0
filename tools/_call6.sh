#!/bin/bash
# full GPU test suite (no -x: list every failure)
python -m pytest tests -m gpu -q -x --deselect tests/test_multi_gpu.py 2>&1 | tail -40 > gpurun_out/c6_pytest.log
python -m pytest tests/test_dropin_models_gpu.py -m gpu -q 2>&1 | tail -60 > gpurun_out/c6_models.log
tail -5 gpurun_out/c6_pytest.log; tail -30 gpurun_out/c6_models.log
