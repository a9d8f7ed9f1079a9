#!/bin/bash
# r02 call H: int8 scan on the TMA-staged kernel: parity tests + throughput
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "int8" > gpurun_out/h_pytest_i8.txt 2>&1
tail -8 gpurun_out/h_pytest_i8.txt
timeout 120 python tools/bench_i8.py > gpurun_out/h_i8.txt 2>&1
timeout 120 python tools/bench_i8.py --rows 4000000 --dim 768 >> gpurun_out/h_i8.txt 2>&1
timeout 120 python tools/bench_i8.py --rows 4000000 --dim 1024 --limit 100 >> gpurun_out/h_i8.txt 2>&1
cat gpurun_out/h_i8.txt
