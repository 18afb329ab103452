#!/usr/bin/env bash
# round 2, capture w (2 GPUs): NCCL exchange behind the C-ABI — multi-GPU tests incl. the plain-C two-thread program, conf5 timing of the three modes
set -u
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q) > gpurun_out/r2w_tests_multi.txt 2>&1; tail -5 gpurun_out/r2w_tests_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/bench_conf5.py > gpurun_out/r2w_conf5_n2.json 2> gpurun_out/r2w_conf5_n2.err; tail -3 gpurun_out/r2w_conf5_n2.err; cat gpurun_out/r2w_conf5_n2.json
