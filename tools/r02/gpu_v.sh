#!/bin/bash
# r02 call V: last check of the cleaned-up session / stream / destroy code on 1 GPU
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_serve_gpu.py tests/test_parity_gpu.py -x -q -k "session or stream or concurrent or trait or config2" > gpurun_out/v_pytest.txt 2>&1
tail -4 gpurun_out/v_pytest.txt
timeout 200 python bench.py --steps 300 --warmup 20 --configs c5 --no-cpu-baseline > gpurun_out/v_bench_n1.json 2> gpurun_out/v_bench_n1.err
tail -2 gpurun_out/v_bench_n1.err; python - <<'PY'
import json
for l in open("gpurun_out/v_bench_n1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["e2e"]["value"], d["parity_ok"], d["session"].get("same_result_as_launch_path"), d["concurrent_callers"]["value"], {c: (v.get("parity_ok"), v.get("ms_per_batch"), v.get("error")) for c, v in d["configs"].items()})
PY
