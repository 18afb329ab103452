#!/usr/bin/env bash
set -u
TAG="${1:-r2_h}"
mkdir -p gpurun_out
summ='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["kernel_ms"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"])'
for m in 0 1 2 4 3 5 6 7; do
  echo "== post_align $m"; python bench.py --no-cpu-baseline --no-config4 --no-full-load --steps 100 --warmup 10 --post-occ 22 --post-align $m 2> gpurun_out/${TAG}_m$m.err | python -c "$summ"
done
