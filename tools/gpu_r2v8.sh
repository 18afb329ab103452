#!/usr/bin/env bash
# round 2, capture v8 (8 GPUs): multi-GPU tests, bench line both arms, and the fused exchange variants for conferences of 16
set -u
bash tools/gpu_multi.sh r2v8 8 tests
for v in "--rs 1" "--rs 0 --tile row" "--rs 0 --tile 16"; do
  tag=$(echo $v | tr -d ' -')
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 tools/bench_conf5.py --conf-size 16 $v > gpurun_out/r2v8_conf5_c16_$tag.json 2> gpurun_out/r2v8_conf5.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2v8_conf5_c16_$tag.json") if l.startswith("{")][-1])
    print("$v", {k[:12]: round(v * 1e3, 2) for k, v in d["ms_per_tick"].items()})
except Exception as ex:
    print("$v", "unreadable", ex)
PY
done
