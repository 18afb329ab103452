#!/usr/bin/env bash
# round 2, capture s2: GPU tests of nsx + AEC after the shape changes, synccheck of the fixed-point suppressor's kernels
set -u
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_nsx.py tests/test_gpu_parity.py -x -q -k "nsx or aec or 32khz") > gpurun_out/r2s2_tests.txt 2>&1; tail -5 gpurun_out/r2s2_tests.txt
SAN_ONLY=nsx SAN_TICKS=3 timeout 1500 compute-sanitizer --tool synccheck python tools/sanitize_small.py > gpurun_out/r2s_synccheck_nsx.txt 2>&1
tail -3 gpurun_out/r2s_synccheck_nsx.txt
timeout 600 python tools/bench_nsx.py --cfgs 7 --align 1 2>&1 | tail -1
