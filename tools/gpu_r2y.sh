#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -k "bus or conference or host or pipelined or g711") > gpurun_out/r2y_tests.txt 2>&1; tail -3 gpurun_out/r2y_tests.txt
for i in 1 2; do timeout 900 python bench.py --steps 200 --no-cpu-baseline --no-config4 --no-full-load --no-offline --no-nsx 2> gpurun_out/r2y_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])"; done
