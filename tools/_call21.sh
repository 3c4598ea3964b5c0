#!/bin/bash
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/c21_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c21_bench_n2.json 2> gpurun_out/c21_bench_n2.err
tail -4 gpurun_out/c21_multi.log; cat gpurun_out/c21_bench_n2.json | cut -c1-700; tail -2 gpurun_out/c21_bench_n2.err
