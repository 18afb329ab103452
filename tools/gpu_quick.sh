#!/usr/bin/env bash
# quick GPU iteration: parity tests, then bench lines for a few variants.  usage: tools/gpu_quick.sh <tag> [VAR=VAL ...]
set -u
TAG="${1:-q}"; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
summ='import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step","kernel_ms")}, "frac=%.4f"%d["roofline"]["frac"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"])'
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "$summ" < gpurun_out/${TAG}_bench.json
for v in "$@"; do
  echo "== $v"; env $v python bench.py --no-cpu-baseline --steps 100 2> gpurun_out/${TAG}_var.err | python -c "$summ"
done
