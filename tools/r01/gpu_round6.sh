#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('n_gpus','value','ms_per_step','gpu_launches')}, 'scan_ms',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],3),'share',round(d['roofline']['step_share'],3),'e2e',round(d['e2e']['value'],1), d['clocks'])" $1; }
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu6.txt
echo "== bench N=1 pdl on/off"; timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/b6_n1.json 2> gpurun_out/b6.err; show gpurun_out/b6_n1.json; tail -2 gpurun_out/b6.err
timeout 600 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline --opt pdl=0 > gpurun_out/b6_n1_nopdl.json 2> gpurun_out/b6.err; show gpurun_out/b6_n1_nopdl.json
echo "== bench N=2 p2p on/off"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 > gpurun_out/b6_n2.json 2> gpurun_out/b6_n2.err; show gpurun_out/b6_n2.json; grep -v OMP gpurun_out/b6_n2.err | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1000 --warmup 10 --opt p2p=0 > gpurun_out/b6_n2_nccl.json 2> gpurun_out/b6_n2.err; show gpurun_out/b6_n2_nccl.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1000 --warmup 10 --opt pdl=0 > gpurun_out/b6_n2_nopdl.json 2> gpurun_out/b6_n2.err; show gpurun_out/b6_n2_nopdl.json
echo "== C4 prefetch distance"; for o in 0 8 16 32 64 128; do timeout 600 python tools/bench_tc.py --opt tc_prefetch=$o --opt tc_target=1024 2>&1 | tail -1 | tee -a gpurun_out/tc_c4_pf.txt; done
echo "== C3"; timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_target=1024 2>&1 | tail -1 | tee gpurun_out/tc_c3_b.txt
timeout 900 python tools/bench_tc.py --rows 10000000 --dim 768 --nq 256 --iters 5 --opt tc_target=1024 --opt tc_prefetch=64 2>&1 | tail -1 | tee -a gpurun_out/tc_c3_b.txt
