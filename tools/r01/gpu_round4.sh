#!/bin/bash
mkdir -p gpurun_out
echo "== C4 per-GPU shape (6.25M x 1024 f16, Q=64, k=100)"; timeout 600 python tools/bench_tc.py 2>&1 | tail -2 | tee gpurun_out/tc_c4.txt
for s in 3 4 6; do timeout 600 python tools/bench_tc.py --opt tc_stages=$s 2>&1 | tail -1 | tee -a gpurun_out/tc_c4.txt; done
echo "== C4 k=10"; timeout 600 python tools/bench_tc.py --k 10 2>&1 | tail -1 | tee -a gpurun_out/tc_c4.txt
echo "== C3 (10M x 768 f16, Q=256, k=100)"; timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 2>&1 | tail -1 | tee gpurun_out/tc_c3.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_max_n=96 2>&1 | tail -1 | tee -a gpurun_out/tc_c3.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_max_n=64 2>&1 | tail -1 | tee -a gpurun_out/tc_c3.txt
echo "== ncu tc"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan -s 4 -c 2 -f -o gpurun_out/tc_full python tools/bench_tc.py --iters 2 > gpurun_out/ncu_tc.log 2>&1; tail -3 gpurun_out/ncu_tc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/tc_launches.csv python tools/bench_tc.py --iters 2 > /dev/null 2>&1; tail -25 gpurun_out/tc_launches.csv | cut -c1-260
