#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
if [ "$N" = "1" ]; then
  timeout 900 python tools/bench_c4.py --iters 5 2>&1 | tail -1 | tee gpurun_out/c4_n1.json
  timeout 900 python tools/bench_c4.py --iters 5 --k 10 2>&1 | tail -1 | tee -a gpurun_out/c4_n1.json
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_c4.py --iters 20 2>&1 | grep -v -E "OMP|\*\*\*|^$" | tail -2 | tee gpurun_out/c4_n$N.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/bench_c4.py --iters 20 --k 10 2>&1 | grep -v -E "OMP|\*\*\*|^$" | tail -1 | tee -a gpurun_out/c4_n$N.json
fi
