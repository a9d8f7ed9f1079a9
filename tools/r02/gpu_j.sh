#!/bin/bash
# r02 call J (2 GPUs): multi-device index tests at HEAD (k up to 1024, tensor path, device I/O, non-SIMD formulas, C++ mirror),
# coalescing test + throughput, C3 after the epilogue latency changes, tensor parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/j_pytest_multi.txt 2>&1
tail -15 gpurun_out/j_pytest_multi.txt
timeout 600 python -m pytest tests/test_tensor_gpu.py tests/test_parity_gpu.py -x -q -k "tensor or concurrent or resolver or paired or tf32" > gpurun_out/j_pytest_tc.txt 2>&1
tail -8 gpurun_out/j_pytest_tc.txt
for o in "" "--opt tc_debug=2"; do
  timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 10 $o >> gpurun_out/j_c3.txt 2>&1
  timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 10 $o >> gpurun_out/j_c4shape.txt 2>&1
done
cat gpurun_out/j_c3.txt gpurun_out/j_c4shape.txt
timeout 300 python tools/bench_concurrent.py > gpurun_out/j_concurrent.txt 2>&1
cat gpurun_out/j_concurrent.txt
