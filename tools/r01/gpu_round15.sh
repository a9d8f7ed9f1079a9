#!/bin/bash
mkdir -p gpurun_out
echo "== trace"; for r in 125000 1000000; do for o in "pdl=1" "pdl=0"; do timeout 300 python tools/trace_steps.py --rows $r --opt $o 2>&1 | tail -6; done; done | tee gpurun_out/trace_steps.txt
echo "== pytest tensor (tf32 + all)"; timeout 1200 python -m pytest tests/test_tensor_gpu.py tests/test_golden_fixtures.py -m gpu -x -q 2>&1 | tail -4
echo "== epilogue split A/B on C3 and C4"; for dbg in 0 4; do timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_debug=$dbg 2>&1 | tail -1; timeout 600 python tools/bench_tc.py --opt tc_debug=$dbg 2>&1 | tail -1; done | tee gpurun_out/tc_ab.txt
