#!/bin/bash
# r02 call M: first run of the resident server kernel (tests under a short timeout: a hang must not eat the box)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_serve_gpu.py -x -q -s > gpurun_out/m_pytest_serve.txt 2>&1
echo "rc=$?" >> gpurun_out/m_pytest_serve.txt
tail -40 gpurun_out/m_pytest_serve.txt
nvidia-smi --query-gpu=name,utilization.gpu --format=csv
