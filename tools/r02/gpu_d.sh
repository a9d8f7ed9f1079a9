#!/bin/bash
# r02 call D: per-launch durations of one tensor batch (new flow) for C4 shape and C3
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/d_c4_launches.csv python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 2 > gpurun_out/d_c4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 12 --csv --log-file gpurun_out/d_c3_launches.csv python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 2 > gpurun_out/d_c3.log 2>&1
