#!/bin/bash
# r02 call L: full 1-GPU test suite at HEAD (rw-lock, group commit, NVTX), exact-vs-tensor crossover for small batches,
# compute-sanitizer memcheck / racecheck / synccheck on small shapes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/l_pytest_gpu.txt 2>&1
tail -5 gpurun_out/l_pytest_gpu.txt
timeout 600 python tools/bench_paths.py > gpurun_out/l_paths.txt 2>&1
cat gpurun_out/l_paths.txt
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "geometry or ties or back_to_back" > gpurun_out/l_san_memcheck_exact.txt 2>&1; echo "memcheck exact rc=$?" >> gpurun_out/l_san_summary.txt
timeout 900 $SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_tensor_gpu.py -x -q -k "paired" > gpurun_out/l_san_memcheck_tensor.txt 2>&1; echo "memcheck tensor(paired) rc=$?" >> gpurun_out/l_san_summary.txt
timeout 900 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "ties or back_to_back" > gpurun_out/l_san_racecheck_exact.txt 2>&1; echo "racecheck exact rc=$?" >> gpurun_out/l_san_summary.txt
timeout 900 $SAN --tool synccheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "ties or back_to_back" > gpurun_out/l_san_synccheck_exact.txt 2>&1; echo "synccheck exact rc=$?" >> gpurun_out/l_san_summary.txt
timeout 900 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_tensor_gpu.py -x -q -k "paired" > gpurun_out/l_san_racecheck_tensor.txt 2>&1; echo "racecheck tensor(paired) rc=$?" >> gpurun_out/l_san_summary.txt
cat gpurun_out/l_san_summary.txt
for f in gpurun_out/l_san_*.txt; do echo "== $f"; tail -6 $f; done
