#!/bin/bash
# r02 call T (2 GPUs): all multi-GPU tests at HEAD + the fixed bench line at N=2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/t_pytest_multi.txt 2>&1
tail -4 gpurun_out/t_pytest_multi.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 500 --warmup 20 --configs none > gpurun_out/t_bench_n2.json 2> gpurun_out/t_bench_n2.err
tail -2 gpurun_out/t_bench_n2.err; grep "^{" gpurun_out/t_bench_n2.json | cut -c1-300
