#!/usr/bin/env bash
# Multi-GPU checks on one box (gpurun --gpus N): cross-GPU bus tests, the headline bench at N ranks (both arms),
# and the config-5 measurement (fused peer kernel vs NCCL).  usage: tools/gpu_multi.sh <tag> <N>
set -u
TAG="${1:-m}"; N="${2:-2}"
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -$((N+2)) > gpurun_out/${TAG}_topo.txt
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29601 bench.py --gpus $N --steps 200 --warmup 250 2> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_n$N.json
$TR --master-port 29602 bench.py --impl reference --gpus $N --steps 40 --warmup 5 2>> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_reference_n$N.json
for cs in 1024 16; do
  $TR --master-port 29603 tools/bench_conf5.py --conf-size $cs 2>> gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_conf5_n${N}_c$cs.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, {k:d.get(k) for k in ("value","ms_per_step","n_gpus","ms_per_tick","impl") if k in d})
PY
tail -3 gpurun_out/${TAG}_bench.err
