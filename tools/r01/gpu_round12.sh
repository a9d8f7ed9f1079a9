#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tensor"; timeout 1200 python -m pytest tests/test_tensor_gpu.py -m gpu -x -q 2>&1 | tail -4
echo "== bench C2 (direct-to-host results)"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/b12_n1.json 2> gpurun_out/b12.err; python -c "
import json;d=json.loads(open('gpurun_out/b12_n1.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step')}, 'scan_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/b12.err
echo "== C4 v1 / pairs"; timeout 600 python tools/bench_tc.py 2>&1 | tail -1 | tee gpurun_out/tc_c4_f.txt
timeout 600 python tools/bench_tc.py --opt tc_kernel=2 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_f.txt
echo "== C3 pairs"; timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 2>&1 | tail -1 | tee gpurun_out/tc_c3_f.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_debug=2 2>&1 | tail -1 | tee -a gpurun_out/tc_c3_f.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 10 --iters 5 2>&1 | tail -1 | tee -a gpurun_out/tc_c3_f.txt
