#!/usr/bin/env bash
# round 2, call E (2 GPUs): multi-GPU parity tests, bench with the conference-bus exchange (conf5) under torchrun
set -u
TAG="${1:-r2_e}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
echo skip-multi
echo skip-parity
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10) > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -3 gpurun_out/${TAG}_bench_n2.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 20 --warmup 5) > gpurun_out/${TAG}_bench_reference_n2.json 2>> gpurun_out/${TAG}_bench_n2.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n2.json", "gpurun_out/${TAG}_bench_reference_n2.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms", "full_load", "conf5")}, d.get("e2e", {}).get("ms_per_step"), d.get("e2e", {}).get("copy_ceiling_ms_per_step"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
