#!/bin/bash
# 8-GPU box: multi-GPU parity tests + strong-scaling bench points
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','gpu_launches')}, 'scan_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],3),'share',round(d['roofline']['step_share'],3),'e2e',round(d['e2e']['value'],1), d['clocks'])" $1; }
nvidia-smi -L | head -8
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2000 --warmup 20 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; show gpurun_out/scale_n$N.json; grep -v -E "OMP|\*\*\*|^$" gpurun_out/scale_n$N.err | tail -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 2000 --warmup 20 --opt p2p=0 > gpurun_out/scale_n8_nccl.json 2> /dev/null; show gpurun_out/scale_n8_nccl.json
