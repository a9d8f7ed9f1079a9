#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu5.txt
echo "== bench (pdl on)"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; python -c "
import json;d=json.load(open('gpurun_out/bench5.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; tail -3 gpurun_out/bench5.err
echo "== bench (pdl off)"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline --opt pdl=0 > gpurun_out/bench5b.json 2> gpurun_out/bench5.err; python -c "
import json;d=json.load(open('gpurun_out/bench5b.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"
echo "== C4 shape"; for o in "tc_target=2048" "tc_target=1024" "tc_target=4096" "tc_l2promo=1" "tc_l2promo=0"; do timeout 600 python tools/bench_tc.py --opt $o 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_b.txt; done
echo "== launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/tc_launches2.csv python tools/bench_tc.py --iters 2 > /dev/null 2>&1; tail -22 gpurun_out/tc_launches2.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
