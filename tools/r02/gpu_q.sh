#!/bin/bash
# r02 call Q (2 GPUs): sessions on a sharded index, bench line at N=1 (short, no batched configs) and N=2 with the session records
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "session or sharded_search" > gpurun_out/q_pytest_multi.txt 2>&1
tail -5 gpurun_out/q_pytest_multi.txt
timeout 600 python bench.py --steps 500 --warmup 20 --configs none > gpurun_out/q_bench_n1.json 2> gpurun_out/q_bench_n1.err
tail -3 gpurun_out/q_bench_n1.err
python - <<'PY'
import json
for f in ["gpurun_out/q_bench_n1.json"]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print({k: d[k] for k in ("value", "ms_per_step", "batch1_transport", "transports", "session", "concurrent_callers", "gpu_launches", "parity_ok")})
            print(d["roofline"]); print(d["e2e"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 500 --warmup 20 --configs none > gpurun_out/q_bench_n2.json 2> gpurun_out/q_bench_n2.err
tail -3 gpurun_out/q_bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/q_bench_n2.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "ms_per_step", "batch1_transport", "transports", "session", "gpu_launches", "parity_ok", "exchange")})
        print(d["roofline"]); print(d["e2e"])
PY
