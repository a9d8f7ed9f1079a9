#!/bin/bash
mkdir -p gpurun_out
echo "== triage C4 main shape (results invalid with debug bits; timing only)"
for dbg in 0 1 2 3; do timeout 600 python tools/bench_tc.py --iters 6 --opt tc_debug=$dbg 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('debug',d['opts'],'ms',d['ms_per_batch'],'GBps',d['GBps_per_pass'],'fallbacks',d['tc_fallbacks'])"; done | tee gpurun_out/tc_triage.txt
echo "== pytest tensor"; timeout 1200 python -m pytest tests/test_tensor_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== pytest parity (int8, flat)"; timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
