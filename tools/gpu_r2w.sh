#!/usr/bin/env bash
# round 2, capture w: G.711-leg host tick, nsx offline across a re-learning
set -u
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_nsx.py -x -q -k "g711_legs or relearning or offline") > gpurun_out/r2w_tests.txt 2>&1; tail -12 gpurun_out/r2w_tests.txt
