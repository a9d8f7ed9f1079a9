#!/bin/bash
# r02 call A: pipeline-step microbenchmark, baseline C3 / C4-shape tensor numbers at HEAD, ncu of both tensor kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt
timeout 120 tools/_bin/tma_stream > gpurun_out/a_tma_stream.txt 2>&1
timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 10 > gpurun_out/a_c4shape.txt 2>&1
timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 10 > gpurun_out/a_c3.txt 2>&1
timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 10 --opt tc_debug=2 > gpurun_out/a_c4shape_noepi.txt 2>&1
timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 10 --opt tc_debug=2 > gpurun_out/a_c3_noepi.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc2_scan_kernel -s 5 -c 1 -o gpurun_out/a_tc2_c3 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 1 > gpurun_out/a_ncu_tc2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 5 -c 1 -o gpurun_out/a_tc_c4 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 1 > gpurun_out/a_ncu_tc.log 2>&1
ls -la gpurun_out
