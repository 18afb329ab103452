#!/usr/bin/env bash
# steady-state `--set full` capture of the AEC kernel (past tick 420) -> gpurun_out/<tag>_aec.ncu-rep
set -u
TAG="${1:-x}"
mkdir -p gpurun_out
bash tools/gpu_aec.sh ${TAG}
ncu --set full --clock-control none --import-source on -k regex:aec_kernel -s 420 -c 1 -o gpurun_out/${TAG}_aec -f \
    python tools/bench_aec.py --steps 30 --warmup 400 --no-ns > /dev/null 2>&1
ls -la gpurun_out | tail -3
