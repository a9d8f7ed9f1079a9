#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu2.txt
echo "== bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
echo "== sweep"; timeout 600 python tools/sweep.py --tiles 8,16,32 --stages 2,4,8 > gpurun_out/sweep.txt 2>&1; tail -30 gpurun_out/sweep.txt
echo "== sweep f16 1024"; timeout 600 python tools/sweep.py --dtype f16 --dim 1024 --rows 2000000 --tiles 16,32 --stages 4,8 --hints 0 > gpurun_out/sweep_f16.txt 2>&1; tail -12 gpurun_out/sweep_f16.txt
echo "== sweep nq"; for q in 2 4; do timeout 300 python tools/sweep.py --nq $q --tiles 16 --stages 4 --hints 0 2>&1 | tail -2; done | tee gpurun_out/sweep_nq.txt
