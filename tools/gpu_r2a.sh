#!/usr/bin/env bash
# round 2, call A: parity tests, both bench arms with the new plumbing, compute-sanitizer logs
set -u
TAG="${1:-r2_a}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -5 gpurun_out/${TAG}_tests.txt
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
(time python bench.py --steps 20 --warmup 5) > gpurun_out/${TAG}_bench_w5.json 2> gpurun_out/${TAG}_bench_w5.err
tail -3 gpurun_out/${TAG}_bench_w5.err
(time python bench.py) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_w5.json", "gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms", "full_load")}, d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("copy_ceiling_ms_per_step"), d.get("e2e", {}).get("bus_only"), d.get("config4", {}).get("kernel_ms"), d.get("cpu_baseline", {}).get("value"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_memcheck.txt 2>&1; tail -4 gpurun_out/${TAG}_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_racecheck.txt 2>&1; tail -4 gpurun_out/${TAG}_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_synccheck.txt 2>&1; tail -4 gpurun_out/${TAG}_synccheck.txt
ls -la gpurun_out | tail -12
