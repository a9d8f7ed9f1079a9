#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','gpu_launches')}, 'scan_ms',round(d['roofline']['kernel_ms'],4),'share',round(d['roofline']['step_share'],3),'e2e',round(d['e2e']['value'],1), d['clocks'])" $1; }
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 2000 --warmup 20 > gpurun_out/scale2_n$N.json 2> gpurun_out/scale2.err; show gpurun_out/scale2_n$N.json; grep -v -E "OMP|\*\*\*|^$" gpurun_out/scale2.err | tail -2
done
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/scale2_n1.json 2>/dev/null; show gpurun_out/scale2_n1.json
