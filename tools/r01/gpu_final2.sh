#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final.txt
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'scan_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e'], d['clocks'])"; tail -2 gpurun_out/bench_final2.err
