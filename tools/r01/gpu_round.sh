#!/bin/bash
# One GPU session: smoke, parity tests, bench, geometry sweep, ncu launch list + full capture of the scan kernel.
# Everything the builder wants back goes to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 1000 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== sweep"; timeout 600 python tools/sweep.py > gpurun_out/sweep.txt 2>&1; tail -40 gpurun_out/sweep.txt
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -15 gpurun_out/launches.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_exact -s 3 -c 2 -f -o gpurun_out/scan_full python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; ls -la gpurun_out/
