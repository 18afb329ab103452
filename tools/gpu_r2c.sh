#!/usr/bin/env bash
# round 2, call C: CTA-cooperative NS kernel v2 (overlapped reducer) — parity, racecheck, shape sweep, ncu capture
set -u
TAG="${1:-r2_c}"
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -5 gpurun_out/${TAG}_tests.txt
summ='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ("value","ms_per_step","kernel_ms")}, "frac=%.4f"%d["roofline"]["frac"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"])'
for c in 0 4 2 6; do
  echo "== ns_cfg $c"; python bench.py --no-cpu-baseline --no-config4 --no-full-load --steps 100 --warmup 10 --ns-cfg $c 2> gpurun_out/${TAG}_cfg$c.err | tee gpurun_out/${TAG}_cfg$c.json | python -c "$summ"
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_racecheck.txt 2>&1; tail -4 gpurun_out/${TAG}_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_memcheck.txt 2>&1; tail -3 gpurun_out/${TAG}_memcheck.txt
ncu --set full --clock-control none --import-source on -k regex:ns_cta_kernel -s 605 -c 1 -o gpurun_out/${TAG}_ns -f \
    python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-config4 --no-full-load > /dev/null 2>&1
ls -la gpurun_out | tail -8
