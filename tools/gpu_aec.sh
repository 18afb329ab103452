#!/usr/bin/env bash
# AEC iteration on the GPU box: the AEC parity tests, then the config-4 measurement of the AEC kernel alone, once per
# VAR=VAL argument (environment knobs of the engine).  usage: tools/gpu_aec.sh <tag> [VAR=VAL ...]
set -u
mkdir -p gpurun_out
TAG="${1:-aec}"; shift || true
python -m pytest tests/test_gpu_parity.py -x -q -k "aec or config4" 2>&1 | tail -4
summ="import json,sys; d=json.loads(sys.stdin.read()); print('aec only', d['kernel_ms']['aec_kernel'], 'frac=%.4f' % d['roofline']['frac'], d['aec_status'])"
python tools/bench_aec.py --no-ns --steps 200 > gpurun_out/${TAG}_aec_only.json 2> gpurun_out/${TAG}_aec.err; python -c "$summ" < gpurun_out/${TAG}_aec_only.json
for v in "$@"; do
  echo "== $v"; env ${v//,/ } python tools/bench_aec.py --no-ns --steps 200 2>> gpurun_out/${TAG}_aec.err | python -c "$summ"
done
