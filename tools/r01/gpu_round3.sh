#!/bin/bash
mkdir -p gpurun_out
echo "== tc debug"; timeout 300 python tools/tc_debug.py 2>&1 | tail -30 | tee gpurun_out/tc_debug.txt
echo "== pytest gpu (parity)"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu3.txt
echo "== pytest gpu (tensor)"; timeout 900 python -m pytest tests/test_tensor_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/pytest_tensor.txt
