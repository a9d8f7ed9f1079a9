#!/bin/bash
# r02 call E: full 1-GPU test suite + the new bench line (all configs)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest_gpu.txt 2>&1
tail -8 gpurun_out/e_pytest_gpu.txt
( time timeout 900 python bench.py --steps 500 --warmup 20 ) > gpurun_out/e_bench_n1.json 2> gpurun_out/e_bench_n1.err
tail -5 gpurun_out/e_bench_n1.err
cat gpurun_out/e_bench_n1.json | head -c 6000
