#!/usr/bin/env bash
set -u
TAG="${1:-r2_j}"
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -4 gpurun_out/${TAG}_tests.txt
python bench.py --no-cpu-baseline --no-config4 --no-full-load --steps 50 --warmup 10 2> gpurun_out/${TAG}_b.err | tee gpurun_out/${TAG}_bench.json | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["kernel_ms"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"], d["value"]); print(d.get("offline"))'
tail -3 gpurun_out/${TAG}_b.err
