#!/bin/bash
mkdir -p gpurun_out
echo "== triage C4 (scan_ms_events = tensor level kernels + selects only)"
for dbg in 0 1 2 3; do timeout 600 python tools/bench_tc.py --iters 4 --opt tc_debug=$dbg 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('debug',d['opts'],'scan_ms',d['scan_ms_events'],'GBps(scan only)',round(d['rows']*d['dim']*2/d['scan_ms_events']/1e6,1),'total_ms',d['ms_per_batch'],'fallbacks',d['tc_fallbacks'])"; done | tee gpurun_out/tc_triage.txt
