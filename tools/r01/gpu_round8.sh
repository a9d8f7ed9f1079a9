#!/bin/bash
mkdir -p gpurun_out
echo "== tc2 debug"; timeout 300 python - <<'PY' 2>&1 | tail -12
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import tc_debug as t
for args in [(2000,64,16,5),(20000,128,16,10),(100000,128,64,10),(100000,768,256,100),(200000,1024,64,100)]:
    t.case(*args, tc_kernel=2)
PY
echo "== pytest tensor + int8"; timeout 1200 python -m pytest tests/test_tensor_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu8.txt
echo "== C4 v1 vs pairs"; timeout 600 python tools/bench_tc.py 2>&1 | tail -1 | tee gpurun_out/tc_c4_d.txt
for st in 0 6 8 10; do timeout 600 python tools/bench_tc.py --opt tc_kernel=2 --opt tc_stages=$st 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_d.txt; done
echo "== C3 v1 vs pairs"; timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_kernel=1 2>&1 | tail -1 | tee gpurun_out/tc_c3_d.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 2>&1 | tail -1 | tee -a gpurun_out/tc_c3_d.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_stages=4 2>&1 | tail -1 | tee -a gpurun_out/tc_c3_d.txt
