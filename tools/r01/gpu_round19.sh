#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','gpu_launches')}, 'scan_ms',round(d['roofline']['kernel_ms'],4),'share',round(d['roofline']['step_share'],3),'e2e',round(d['e2e']['value'],1))" $1; }
echo "== multi-gpu tests"; timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
for o in 2 1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$o bench.py --gpus 2 --steps 2000 --warmup 20 --opt pdl=$o > gpurun_out/b19_n2_pdl$o.json 2> gpurun_out/b19.err; show gpurun_out/b19_n2_pdl$o.json; grep -v -E "OMP|\*\*\*|^$" gpurun_out/b19.err | tail -2
done
