#!/usr/bin/env bash
# fixed-point suppressor on the GPU box: parity tests, kernel timing of every launch shape, ncu launch list + one full capture
# usage: tools/gpu_nsx.sh <tag>
set -u
TAG="${1:-nsx}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_nsx.py -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -15 gpurun_out/${TAG}_tests.txt
timeout 600 python tools/bench_nsx.py --float-core > gpurun_out/${TAG}_bench_nsx.jsonl 2> gpurun_out/${TAG}_bench_nsx.err; cat gpurun_out/${TAG}_bench_nsx.jsonl; tail -3 gpurun_out/${TAG}_bench_nsx.err
timeout 600 python tools/bench_nsx.py --freq 8000 --cfgs 0,1,2 --align 1 > gpurun_out/${TAG}_bench_nsx8k.jsonl 2>> gpurun_out/${TAG}_bench_nsx.err; cat gpurun_out/${TAG}_bench_nsx8k.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsx_kernel -s 615 -c 1 -o gpurun_out/${TAG}_nsx -f \
    python tools/bench_nsx.py --cfgs 0 --align 1 --steps 10 > /dev/null 2>&1
ls -la gpurun_out | tail -6
