#!/bin/bash
# r02 call N: resident session vs launch-per-query timings (C2 and its per-GPU shard sizes)
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_serve.py --chunks 0 2 4 8 > gpurun_out/n_serve.txt 2>&1
cat gpurun_out/n_serve.txt
