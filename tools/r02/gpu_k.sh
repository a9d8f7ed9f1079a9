#!/bin/bash
# r02 call K: interleaved A/B of the epilogue latency changes (old kernels in tools/_bin/libcgvec_old.so) on one box
set -x
mkdir -p gpurun_out
for rep in 1 2 3; do
  for lib in old new; do
    if [ $lib = old ]; then export CGVEC_AB_LIB=$PWD/tools/_bin/libcgvec_old.so; else unset CGVEC_AB_LIB; fi
    echo "== $lib rep $rep" >> gpurun_out/k_ab.txt
    timeout 200 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --k 100 --iters 20 >> gpurun_out/k_ab.txt 2>&1
    timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 100 --iters 20 >> gpurun_out/k_ab.txt 2>&1
    timeout 200 python tools/bench_tc.py --rows 6250000 --dim 1024 --nq 64 --k 10 --iters 20 >> gpurun_out/k_ab.txt 2>&1
  done
done
unset CGVEC_AB_LIB
timeout 200 python tools/bench_tc.py --rows 12500000 --dim 384 --nq 256 --k 10 --iters 10 --dtype f32 >> gpurun_out/k_c5shape.txt 2>&1
grep -v "^==" gpurun_out/k_ab.txt | cut -c1-140
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "concurrent or trait or flat or int8 or upsert" > gpurun_out/k_pytest.txt 2>&1
tail -5 gpurun_out/k_pytest.txt
