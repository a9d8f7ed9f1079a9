#!/bin/bash
# r02 call S: full 1-GPU test suite at HEAD + smoke + a short C2 bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s_pytest_gpu.txt 2>&1
tail -6 gpurun_out/s_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.txt 2>&1; cat gpurun_out/s_smoke.txt
timeout 300 python bench.py --steps 500 --warmup 20 --configs none --no-cpu-baseline > gpurun_out/s_bench_n1.json 2> gpurun_out/s_bench_n1.err
tail -2 gpurun_out/s_bench_n1.err; cut -c1-400 gpurun_out/s_bench_n1.json
