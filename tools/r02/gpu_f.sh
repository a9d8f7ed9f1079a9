#!/bin/bash
# r02 call F (2 GPUs): multi-GPU parity tests at HEAD + bench at N=2
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_parity_gpu.py::test_query_stream_double_buffered_batches -x -q > gpurun_out/f_pytest_multi.txt 2>&1
tail -15 gpurun_out/f_pytest_multi.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 500 --warmup 20 ) > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err
tail -5 gpurun_out/f_bench_n2.err
head -c 1500 gpurun_out/f_bench_n2.json
