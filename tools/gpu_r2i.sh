#!/usr/bin/env bash
set -u
TAG="${1:-r2_i}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aec or post_kernel or full_chain" 2>&1 | tail -3
for a in 0 1; do echo "== aec_align $a"; python tools/bench_aec.py --steps 100 --warmup 420 --aec-align $a 2>gpurun_out/${TAG}_aec$a.err | tee gpurun_out/${TAG}_aec$a.json | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["kernel_ms"], d["roofline"]["frac"])'; done
python bench.py --no-cpu-baseline --no-config4 --no-full-load --steps 100 --warmup 10 2> gpurun_out/${TAG}_b.err | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["kernel_ms"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"], d["value"])'
