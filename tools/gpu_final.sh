#!/bin/bash
# Round-end validation on one GPU: smoke, full GPU test-suite, bench (with CPU baseline), reference arm, ncu launch list.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_final.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 2600 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_final.json 2>/dev/null; tail -c 700 gpurun_out/bench_ref_final.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; tail -6 gpurun_out/launches_final.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
