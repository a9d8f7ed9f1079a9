#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tensor"; timeout 1200 python -m pytest tests/test_tensor_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== C4 v1 / pairs"; timeout 600 python tools/bench_tc.py 2>&1 | tail -1 | tee gpurun_out/tc_c4_g.txt
timeout 600 python tools/bench_tc.py --opt tc_kernel=2 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_g.txt
echo "== C3 pairs"; timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 2>&1 | tail -1 | tee gpurun_out/tc_c3_g.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_stages=4 2>&1 | tail -1 | tee -a gpurun_out/tc_c3_g.txt
echo "== ncu C3 main level"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc2_scan -s 9 -c 1 -f -o gpurun_out/tc2_main python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 2 > gpurun_out/ncu_tc3.log 2>&1; tail -2 gpurun_out/ncu_tc3.log | cut -c1-200
