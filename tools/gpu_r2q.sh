#!/usr/bin/env bash
# round 2, capture q: nsx tests incl. the 32 kHz engine, float warp-kernel shapes with big aligned CTAs, nsx at the default shape
set -u
TAG=r2q
mkdir -p gpurun_out
echo skip-tests
timeout 900 python tools/bench_nsx.py --float-core --float-cfgs=-1,6,8,9,10,11 --cfgs 7 --align 1 > gpurun_out/${TAG}_shapes.jsonl 2> gpurun_out/${TAG}_shapes.err; cat gpurun_out/${TAG}_shapes.jsonl; tail -3 gpurun_out/${TAG}_shapes.err
timeout 900 python tools/bench_nsx.py --freq 8000 --float-core --float-cfgs=-1,8,9,10 --cfgs 7 --align 1 > gpurun_out/${TAG}_shapes8k.jsonl 2>> gpurun_out/${TAG}_shapes.err; cat gpurun_out/${TAG}_shapes8k.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsx_kernel -s 615 -c 1 -o gpurun_out/${TAG}_nsx -f \
    python tools/bench_nsx.py --cfgs 7 --align 1 --steps 10 > /dev/null 2>&1
ls -la gpurun_out | tail -5
