#!/bin/bash
# r02 call C: selector flow — parity tests, then C3 / C4-shape timings (new flow vs legacy ranges, S0 sweep)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tensor_gpu.py -x -q > gpurun_out/c_pytest_tensor.txt 2>&1
tail -15 gpurun_out/c_pytest_tensor.txt
for o in "" "--opt tc_flow=1" "--opt tc_first=2048" "--opt tc_first=8192" "--opt tc_kernel=2" "--opt tc_debug=2"; do
  timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 10 $o >> gpurun_out/c_c4shape.txt 2>&1
done
for o in "" "--opt tc_flow=1" "--opt tc_first=2048" "--opt tc_first=8192" "--opt tc_debug=2"; do
  timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 10 $o >> gpurun_out/c_c3.txt 2>&1
done
timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 10 --iters 10 >> gpurun_out/c_c4shape.txt 2>&1
timeout 200 python tools/bench_tc.py --rows 12500000 --dim 384 --nq 256 --k 10 --iters 5 --dtype f32 >> gpurun_out/c_c5shape.txt 2>&1
cat gpurun_out/c_c4shape.txt gpurun_out/c_c3.txt gpurun_out/c_c5shape.txt
