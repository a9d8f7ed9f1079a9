#!/bin/bash
# r02 call I: C3 triage (does re-streaming the query block through L2 bound K2b?), ncu --set full of both tensor kernels at HEAD
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/i_gpu.txt
for o in "" "--opt tc_debug=2" "--opt tc_debug=4" "--opt tc_debug=6" "--opt tc_debug=1" "--opt tc_debug=3" "--opt tc_debug=7"; do
  timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 10 $o >> gpurun_out/i_c3.txt 2>&1
done
cat gpurun_out/i_c3.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc2_scan_kernel -s 2 -c 2 -o gpurun_out/i_tc2_c3 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 1 > gpurun_out/i_ncu_tc2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_scan_kernel -s 2 -c 2 -o gpurun_out/i_tc_c4 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 1 > gpurun_out/i_ncu_tc.log 2>&1
ls -la gpurun_out | tail -5
