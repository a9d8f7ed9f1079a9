#!/bin/bash
# r02 call W (2 GPUs): sessions on a sharded index + the bench line at N=2, at the final commit
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -k "session or single_process_multi_device_index" > gpurun_out/w_pytest_multi.txt 2>&1
tail -3 gpurun_out/w_pytest_multi.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 300 --warmup 20 --configs c4 > gpurun_out/w_bench_n2.json 2> gpurun_out/w_bench_n2.err
tail -2 gpurun_out/w_bench_n2.err | cut -c1-200
python - <<'PY'
import json
for l in open("gpurun_out/w_bench_n2.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["e2e"]["value"], d["parity_ok"], d["batch1_transport"], d["transports"], d["session"].get("same_result_as_launch_path"))
        for c, v in d["configs"].items():
            print(c, [(r["k"], round(r["ms_per_batch"], 3), r["parity_ok"], r["tc_fallbacks"], r.get("exchange")) for r in v.get("runs", [])], v.get("error"))
PY
