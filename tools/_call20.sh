#!/bin/bash
# final validation of the round: every GPU test, smoke, the bench line, the launch list of the bench under ncu
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c20_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c20_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/c20_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c20_ncu_bench.log 2>&1
tail -6 gpurun_out/c20_pytest.log; tail -2 gpurun_out/c20_smoke.log; cat gpurun_out/c20_bench.json; tail -2 gpurun_out/c20_bench.err
