#!/usr/bin/env bash
# round 2, call G: full GPU test suite + bench with the aligned AGC+VAD shape as default
set -u
TAG="${1:-r2_g}"
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -5 gpurun_out/${TAG}_tests.txt
summ='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ("value","ms_per_step","kernel_ms")}, "e2e_ms=%.3f"%d["e2e"]["ms_per_step"])'
for c in 0 3; do
  echo "== post_occ $c"; python bench.py --no-cpu-baseline --no-config4 --no-full-load --steps 100 --warmup 10 --post-occ $c 2> gpurun_out/${TAG}_occ$c.err | tee gpurun_out/${TAG}_occ$c.json | python -c "$summ"
done
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 605 -c 1 -o gpurun_out/${TAG}_post -f \
    python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-config4 --no-full-load > /dev/null 2>&1
ls -la gpurun_out | tail -4
