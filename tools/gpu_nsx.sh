#!/usr/bin/env bash
# fixed-point suppressor on the GPU box: parity tests, kernel timing of launch shapes x alignment masks, one ncu full capture
# usage: tools/gpu_nsx.sh <tag> [cfgs] [masks] [ncu_cfg] [ncu_mask]
set -u
TAG="${1:-nsx}"; CFGS="${2:-0,2,4,6,7,8,9}"; MASKS="${3:-1,255}"; NCFG="${4:-4}"; NMASK="${5:-1}"
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_nsx.py -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -15 gpurun_out/${TAG}_tests.txt
timeout 900 python tools/bench_nsx.py --float-core --cfgs $CFGS --align $MASKS > gpurun_out/${TAG}_bench_nsx.jsonl 2> gpurun_out/${TAG}_bench_nsx.err; cat gpurun_out/${TAG}_bench_nsx.jsonl; tail -3 gpurun_out/${TAG}_bench_nsx.err
timeout 600 python tools/bench_nsx.py --freq 8000 --cfgs $CFGS --align $NMASK > gpurun_out/${TAG}_bench_nsx8k.jsonl 2>> gpurun_out/${TAG}_bench_nsx.err; cat gpurun_out/${TAG}_bench_nsx8k.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsx_kernel -s 615 -c 1 -o gpurun_out/${TAG}_nsx -f \
    python tools/bench_nsx.py --cfgs $NCFG --align $NMASK --steps 10 > /dev/null 2>&1
ls -la gpurun_out | tail -6
