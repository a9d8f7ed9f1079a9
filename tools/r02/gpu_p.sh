#!/bin/bash
# r02 call P: session streaming rate vs L2 hint
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_serve.py --chunks 0 --steps 1000 --rows 1000000 125000 1000000 > gpurun_out/p_serve.txt 2>&1
cat gpurun_out/p_serve.txt
timeout 300 python -m pytest tests/test_serve_gpu.py -x -q > gpurun_out/p_pytest_serve.txt 2>&1; tail -3 gpurun_out/p_pytest_serve.txt
