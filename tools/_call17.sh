#!/bin/bash
timeout 300 python tools/cond_profile.py 50000 4 > gpurun_out/c17_cond_profile.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cond_node_kernel|branch_kernel|eigen_qrev_kernel|pt3_kernel" --launch-skip 40 --launch-count 8 -o gpurun_out/cond python tools/cond_profile.py 50000 4 > gpurun_out/c17_ncu.log 2>&1
cat gpurun_out/c17_cond_profile.log; tail -5 gpurun_out/c17_ncu.log
