#!/usr/bin/env bash
# Config-5 iteration on N GPUs of one box (gpurun --gpus N): cross-GPU bus tests, then the fused peer kernel against
# bus_sum -> NCCL all-reduce -> nminus1 for conferences of 1024 and 16, the latter with both tile shapes.
# usage: tools/gpu_conf5.sh <tag> <N>
set -u
TAG="${1:-c5}"; N="${2:-2}"
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show='import json,sys; d=json.loads(sys.stdin.read()); print({k.split(" ")[0]: round(v*1e3,1) for k,v in d["ms_per_tick"].items()}, "us")'
for cs in 1024 16; do
  $TR --master-port 29603 tools/bench_conf5.py --conf-size $cs 2>> gpurun_out/${TAG}.err | tail -1 > gpurun_out/${TAG}_conf5_n${N}_c$cs.json
  echo "conf $cs:"; python -c "$show" < gpurun_out/${TAG}_conf5_n${N}_c$cs.json
done
echo "conf 16, tile 16:"; WMIXB_PEER_TILE=16 $TR --master-port 29604 tools/bench_conf5.py --conf-size 16 2>> gpurun_out/${TAG}.err | tail -1 | python -c "$show"
tail -2 gpurun_out/${TAG}.err
