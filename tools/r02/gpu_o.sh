#!/bin/bash
# r02 call O: server kernel v2 (helper warp, static tiles): tests, then timings
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_serve_gpu.py -x -q -s > gpurun_out/o_pytest_serve.txt 2>&1
echo "rc=$?" >> gpurun_out/o_pytest_serve.txt
tail -15 gpurun_out/o_pytest_serve.txt
timeout 600 python tools/bench_serve.py > gpurun_out/o_serve.txt 2>&1
cat gpurun_out/o_serve.txt
