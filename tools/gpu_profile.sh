#!/usr/bin/env bash
# Run on the GPU box through gpurun: parity tests, the bench line (both arms, the driver's flags and the defaults), the ncu
# launch list, steady-state `--set full` captures of the NS, post (AGC+VAD) and AEC kernels.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG="${1:-x}"; shift || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
(time python -m pytest tests -m gpu -x -q "$@") > gpurun_out/${TAG}_tests.txt 2>&1; tail -4 gpurun_out/${TAG}_tests.txt
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_w5.json", "gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms")}, "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
# PRIME (600) + warm-up (10) ticks x 3 launches, plus a handful of init launches
ncu --metrics gpu__time_duration.sum --clock-control none -s 1840 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-config4 --no-full-load --no-offline --no-nsx > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ns_cta_kernel -s 605 -c 1 -o gpurun_out/${TAG}_ns -f \
    python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-config4 --no-full-load --no-offline --no-nsx > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:post_kernel -s 605 -c 1 -o gpurun_out/${TAG}_post -f \
    python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-config4 --no-full-load --no-offline --no-nsx > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:aec_kernel -s 420 -c 1 -o gpurun_out/${TAG}_aec -f \
    python tools/bench_aec.py --steps 30 --warmup 400 --no-ns > /dev/null 2>&1
# the staged (persistent offline) NS kernel: DRAM bytes per frame against the tick kernel
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:ns_cta_kernel -s 600 -c 4 --csv \
    --log-file gpurun_out/${TAG}_offline_dram.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config4 --no-full-load > /dev/null 2>&1
# the reports are ~37 MB each and gpurun brings back at most 64 MiB: summarise them here, keep only the NS one
for k in ns post aec; do
  python tools/ncu_summary.py gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_${k}_ncu.md 2>/dev/null
done
rm -f gpurun_out/${TAG}_post.ncu-rep gpurun_out/${TAG}_aec.ncu-rep
ls -la gpurun_out | tail -12
