#!/usr/bin/env bash
# round 2, call F: AGC+VAD kernel shapes (single-wave 704-thread CTA vs 3 x 128)
set -u
TAG="${1:-r2_f}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integer_stages or full_chain or host_buffer or vad_20ms" 2>&1 | tail -3
summ='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ("value","ms_per_step","kernel_ms")}, "e2e_ms=%.3f"%d["e2e"]["ms_per_step"])'
for c in 3 22; do
  echo "== post_occ $c"; python bench.py --no-cpu-baseline --no-config4 --no-full-load --steps 100 --warmup 10 --post-occ $c 2> gpurun_out/${TAG}_occ$c.err | tee gpurun_out/${TAG}_occ$c.json | python -c "$summ"
done
