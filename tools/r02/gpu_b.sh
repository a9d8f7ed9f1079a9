#!/bin/bash
# r02 call B: tensor-path parity tests on the rewritten kernels, then C3 / C4-shape timings over kernel / kbs variants
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tensor_gpu.py -x -q > gpurun_out/b_pytest_tensor.txt 2>&1
tail -5 gpurun_out/b_pytest_tensor.txt
for o in "" "--opt tc_kernel=2" "--opt tc_kernel=2 --opt tc_kbs=4" "--opt tc_kbs=1" "--opt tc_kbs=4" "--opt tc_debug=2" "--opt tc_kernel=2 --opt tc_debug=2"; do
  timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 10 $o >> gpurun_out/b_c4shape.txt 2>&1
done
for o in "" "--opt tc_kbs=1" "--opt tc_kbs=3" "--opt tc_debug=2" "--opt tc_kbs=3 --opt tc_debug=2"; do
  timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 10 $o >> gpurun_out/b_c3.txt 2>&1
done
cat gpurun_out/b_c4shape.txt gpurun_out/b_c3.txt
