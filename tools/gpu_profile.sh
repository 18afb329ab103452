#!/usr/bin/env bash
# Run on the GPU box through gpurun: parity tests, the bench line (both arms), the ncu launch list, steady-state
# `--set full` captures of the NS, post (AGC+VAD) and AEC kernels, and the config-4 measurement.
# usage: tools/gpu_profile.sh <tag> [pytest-args]
set -u
TAG="${1:-x}"; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -8
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 40 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kernel_ms','roofline','e2e','cpu_baseline')})"
python tools/bench_aec.py > gpurun_out/${TAG}_aec.json 2> gpurun_out/${TAG}_aec.err; cat gpurun_out/${TAG}_aec.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 750 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 30 --warmup 250 --no-cpu-baseline > /dev/null 2>&1
# frame 260+: past start-up (50), gain map on (200)
ncu --set full --clock-control none --import-source on -k regex:ns_kernel -s 260 -c 1 -o gpurun_out/${TAG}_ns -f \
    python bench.py --steps 20 --warmup 250 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 260 -c 1 -o gpurun_out/${TAG}_post -f \
    python bench.py --steps 20 --warmup 250 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:aec_kernel -s 420 -c 1 -o gpurun_out/${TAG}_aec -f \
    python tools/bench_aec.py --steps 30 --warmup 400 --no-ns > /dev/null 2>&1
ls -la gpurun_out | tail -12
