#!/usr/bin/env bash
# N-GPU run on one box: bench line with the conference-bus exchange (conf5), reference arm, optional multi-GPU parity tests.
# usage: tools/gpu_multi.sh <tag> <N> [tests]
set -u
TAG="${1:-m}"; N="${2:-2}"; WITH_TESTS="${3:-}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
if [ -n "$WITH_TESTS" ]; then (time python -m pytest tests/test_multi_gpu.py -m gpu -x -q) > gpurun_out/${TAG}_tests_multi.txt 2>&1; tail -3 gpurun_out/${TAG}_tests_multi.txt; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -2 gpurun_out/${TAG}_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_n$N.json 2>> gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n$N.json", "gpurun_out/${TAG}_bench_reference_n$N.json"):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        c5 = d.get("conf5") or {}
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"), "ceiling", d.get("e2e", {}).get("copy_ceiling_ms_per_step"),
              "bus_only", (d.get("e2e", {}).get("bus_only") or {}).get("ms_per_step"), {k: (v["peer_us"], v["nccl_us"], v.get("nccl_c_us"), v["parity_ok"]) for k, v in c5.items()}, (d.get("full_load") or {}).get("ms_per_tick"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
