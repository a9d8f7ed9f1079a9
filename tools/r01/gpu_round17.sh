#!/bin/bash
mkdir -p gpurun_out
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -12
echo "== C4 small sharded"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/bench_c4.py --rows 12500000 --iters 10 2>&1 | grep -v -E "OMP|\*\*\*|^$" | tail -3
