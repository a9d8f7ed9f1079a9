#!/bin/bash
# r02 call G: full 1-GPU test suite (new tests: resolver dense pass, int8 ties, stream, candidate stage), int8 throughput, C3 k/flow A/B
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest_gpu.txt 2>&1
tail -15 gpurun_out/g_pytest_gpu.txt
timeout 120 python tools/bench_i8.py > gpurun_out/g_i8.txt 2>&1
timeout 120 python tools/bench_i8.py --rows 4000000 --dim 768 >> gpurun_out/g_i8.txt 2>&1
cat gpurun_out/g_i8.txt
for o in "--k 10" "--k 10 --opt tc_flow=1" "--k 100" "--k 100 --opt tc_flow=1"; do
  timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --iters 10 $o >> gpurun_out/g_c4k.txt 2>&1
done
cat gpurun_out/g_c4k.txt
