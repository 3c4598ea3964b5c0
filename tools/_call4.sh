mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prune4v2 -s 2 -c 1 -f -o gpurun_out/p4v2 python tools/quick_bench.py cfg5 120000 > gpurun_out/c4.log 2>&1
tail -2 gpurun_out/c4.log
