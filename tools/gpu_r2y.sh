#!/usr/bin/env bash
# round 2, capture y2: ncu --set full of post_kernel after the packed minimum tracker
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 605 -c 1 -o gpurun_out/r2y_post -f \
    python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-config4 --no-full-load --no-offline --no-nsx > /dev/null 2>&1
ls -la gpurun_out/r2y_post.ncu-rep
