#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu11.txt
echo "== bench C2"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/b11_n1.json 2> gpurun_out/b11.err; python -c "
import json;d=json.loads(open('gpurun_out/b11_n1.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step')}, 'scan_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/b11.err
echo "== sweep K1"; timeout 300 python tools/sweep.py --tiles 16,32 --stages 2,4 --hints 0 2>&1 | tail -5
echo "== C4 v1 / pairs"; timeout 600 python tools/bench_tc.py 2>&1 | tail -1 | tee gpurun_out/tc_c4_e.txt
timeout 600 python tools/bench_tc.py --opt tc_kernel=2 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_e.txt
echo "== C3 pairs"; timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 2>&1 | tail -1 | tee gpurun_out/tc_c3_e.txt
echo "== int8 quick"; timeout 300 python - <<'PY' 2>&1 | tail -3
import sys,time; sys.path.insert(0,'.')
import numpy as np, __graft_entry__ as ge
cg=ge.load_package(); import torch
ix=cg.Index(768); ix.fill_synthetic(1_000_000, 1, True); ix.quantize_i8()
q=np.random.default_rng(0).standard_normal(768).astype(np.float32)*0.05
for _ in range(3): ix.search_optimized(q,10)
t=time.perf_counter(); n=50
for _ in range(n): ix.search_optimized(q,10)
dt=(time.perf_counter()-t)/n
print(f"int8 1M x 768 e2e {dt*1e3:.3f} ms/query  -> {768e6/dt/1e9:.0f} GB/s of codes")
PY
