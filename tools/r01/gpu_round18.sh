#!/bin/bash
mkdir -p gpurun_out
echo "== pdl chain test"; timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "back_to_back" 2>&1 | tail -4
echo "== trace"; for r in 125000 1000000; do for o in "pdl=1" "pdl=2"; do timeout 200 python tools/trace_steps.py --rows $r --opt $o 2>&1 | tail -6; done; done | tee gpurun_out/trace_steps2.txt
echo "== bench N=1 pdl=2"; timeout 400 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline --opt pdl=2 > gpurun_out/b18.json 2> gpurun_out/b18.err; python -c "
import json;d=json.loads(open('gpurun_out/b18.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step')}, 'scan_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"; tail -2 gpurun_out/b18.err
