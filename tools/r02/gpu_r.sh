#!/bin/bash
# r02 call R (8 GPUs): the full bench line at N=8 (C2 headline with both batch-1 transports, C3/C4/C5 with parity), a second
# short C2-only line for run-to-run spread, and the multi-GPU session / skew tests on 4 of the GPUs
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > gpurun_out/r_gpus.txt
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2000 --warmup 20 > gpurun_out/r_bench_n8.json 2> gpurun_out/r_bench_n8.err
echo "bench rc=$?"
tail -3 gpurun_out/r_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 2000 --warmup 20 --configs none > gpurun_out/r_bench_n8_c2only.json 2> gpurun_out/r_bench_n8_c2only.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 2000 --warmup 20 --configs none > gpurun_out/r_bench_n4_c2only.json 2> gpurun_out/r_bench_n4_c2only.err
timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 20 --configs none --no-cpu-baseline > gpurun_out/r_bench_n1_c2only.json 2> gpurun_out/r_bench_n1_c2only.err
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -k "session or tiny or uneven" > gpurun_out/r_pytest_multi.txt 2>&1
tail -4 gpurun_out/r_pytest_multi.txt
python - <<'PY'
import json
for f in ["gpurun_out/r_bench_n8.json", "gpurun_out/r_bench_n8_c2only.json", "gpurun_out/r_bench_n4_c2only.json", "gpurun_out/r_bench_n1_c2only.json"]:
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l)
                print(f, {k: d.get(k) for k in ("value", "ms_per_step", "batch1_transport", "transports", "parity_ok", "exchange")}, d["e2e"]["value"], d["roofline"]["frac"])
                for c, v in (d.get("configs") or {}).items():
                    for r in v.get("runs", []) if isinstance(v, dict) else []:
                        print("   ", c, r.get("k"), r.get("ms_per_batch"), r.get("value"), r.get("parity_ok"), r.get("tc_fallbacks"), r.get("exchange"))
                    if isinstance(v, dict) and "error" in v: print("   ", c, v["error"])
    except Exception as e:
        print(f, "ERR", e)
PY
