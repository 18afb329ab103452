#!/usr/bin/env bash
# Run on the GPU box through gpurun: parity tests, the bench line, the ncu launch list and a
# steady-state `--set full` capture of the NS kernel.  usage: tools/gpu_profile.sh <tag> [pytest-args]
set -u
TAG="${1:-x}"; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -8
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kernel_ms','roofline','e2e')})"
ncu --metrics gpu__time_duration.sum --clock-control none -s 750 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 30 --warmup 250 --no-cpu-baseline > /dev/null 2>&1
# frame 260+: past start-up (50), gain map on (200)
ncu --set full --clock-control none --import-source on -k regex:ns_kernel -s 260 -c 2 -o gpurun_out/${TAG}_ns -f \
    python bench.py --steps 20 --warmup 250 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
