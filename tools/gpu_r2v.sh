#!/usr/bin/env bash
# round 2, capture v: schedules of the pipelined host tick — PCM read-out on streams of its own, AGC+VAD once per tick, chunk counts
set -u
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "pipelined or host") > gpurun_out/r2v_tests.txt 2>&1; tail -4 gpurun_out/r2v_tests.txt
summ='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print("ms_step %.4f e2e_ms %.4f (ceiling %.4f) bus_only %.4f (ceiling %.4f) sync %.4f" % (d["ms_per_step"], e["ms_per_step"], e["copy_ceiling_ms_per_step"], e["bus_only"]["ms_per_step"], e["bus_only"]["copy_ceiling_ms_per_step"], e["sync_call_ms_per_step"]))'
B="timeout 900 python bench.py --steps 100 --no-cpu-baseline --no-config4 --no-full-load --no-offline --no-nsx"
for v in "--host-split-d2h 0" "--host-split-d2h 1" "--host-split-d2h 1 --host-chunks 3" "--host-split-d2h 1 --host-chunks 5" "--host-split-d2h 1 --host-chunks 6" "--host-split-d2h 1 --host-chunks 8" "--host-split-d2h 0" "--host-split-d2h 1"; do
  echo "== $v"; $B $v 2> gpurun_out/r2v_bench.err | python -c "$summ"
done
