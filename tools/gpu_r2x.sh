#!/usr/bin/env bash
# round 2, capture x: the reducer's part 1b (Nyquist SNR + two features) runs behind the workers' segment 2.
# NS parity subset, racecheck / synccheck / memcheck of the NS kernels, bench line.
set -u
TAG="${1:-r2x}"
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ns or offline or wav_config1 or config2 or handle or stereo") > gpurun_out/${TAG}_tests.txt 2>&1; tail -4 gpurun_out/${TAG}_tests.txt
for tool in racecheck synccheck memcheck; do
  SAN_ONLY=ns SAN_TICKS=3 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_${tool}_ns.txt 2>&1; tail -3 gpurun_out/${TAG}_${tool}_ns.txt
done
summ='import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step","kernel_ms")}, "frac=%.4f"%d["roofline"]["frac"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"], "offline", (d.get("offline") or {}))'
timeout 900 python bench.py --no-cpu-baseline --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "$summ" < gpurun_out/${TAG}_bench.json || tail -5 gpurun_out/${TAG}_bench.err
