#!/usr/bin/env bash
# round 2, capture u: packet VAD kernels re-staged through shared memory (vad20 / vad32 / record tick / 32 kHz engine), nsx record tick
set -u
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_nsx.py tests/test_gpu_parity.py -x -q -k "record or vad_20ms or 32khz or nsx") > gpurun_out/r2u_tests.txt 2>&1; tail -5 gpurun_out/r2u_tests.txt
