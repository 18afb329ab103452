#!/usr/bin/env bash
# fixed-point suppressor on the GPU box: parity tests, kernel timing of launch shapes, optionally one ncu full capture
# usage: tools/gpu_nsx.sh <tag> [cfgs] [ncu_cfg|none]
set -u
TAG="${1:-nsx}"; CFGS="${2:-7,6,8}"; NCFG="${3:-none}"
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_nsx.py -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -4 gpurun_out/${TAG}_tests.txt
timeout 900 python tools/bench_nsx.py --cfgs $CFGS --align 1 > gpurun_out/${TAG}_bench_nsx.jsonl 2> gpurun_out/${TAG}_bench_nsx.err; cat gpurun_out/${TAG}_bench_nsx.jsonl; tail -3 gpurun_out/${TAG}_bench_nsx.err
timeout 600 python tools/bench_nsx.py --freq 8000 --cfgs $CFGS --align 1 > gpurun_out/${TAG}_bench_nsx8k.jsonl 2>> gpurun_out/${TAG}_bench_nsx.err; cat gpurun_out/${TAG}_bench_nsx8k.jsonl
if [ "$NCFG" != "none" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nsx_kernel -s 615 -c 1 -o gpurun_out/${TAG}_nsx -f \
      python tools/bench_nsx.py --cfgs $NCFG --align 1 --steps 10 > /dev/null 2>&1
fi
