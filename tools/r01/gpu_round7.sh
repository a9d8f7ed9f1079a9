#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu7.txt
echo "== C4 level sizing"; for o in "tc_target=1024" "tc_first=8192 --opt tc_target=1024" "tc_first=8192 --opt tc_target=2048" "tc_first=8192 --opt tc_target=4096" "tc_first=4096 --opt tc_target=2048"; do timeout 600 python tools/bench_tc.py --opt $o 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_c.txt; done
echo "== C4 k=10"; timeout 600 python tools/bench_tc.py --k 10 --opt tc_first=8192 --opt tc_target=2048 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_c.txt
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/tc_launches3.csv python tools/bench_tc.py --iters 2 --opt tc_first=8192 --opt tc_target=2048 > /dev/null 2>&1; tail -16 gpurun_out/tc_launches3.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100
echo "== ncu main level"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_scan -s 11 -c 1 -f -o gpurun_out/tc_main python tools/bench_tc.py --iters 2 --opt tc_first=8192 --opt tc_target=2048 > gpurun_out/ncu_tc2.log 2>&1; tail -2 gpurun_out/ncu_tc2.log | cut -c1-200
