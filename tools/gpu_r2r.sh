#!/usr/bin/env bash
# round 2, capture r: the whole GPU test suite, smoke, and the bench line of both arms as the driver runs them
set -u
TAG=r2r
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_tests.txt 2>&1; tail -6 gpurun_out/${TAG}_tests.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench.err
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "kernel_ms")}, "frac", d.get("roofline", {}).get("frac"), "e2e", d.get("e2e", {}).get("value"))
        for k in ("nsx", "offline", "config4", "full_load"):
            if k in d: print("  ", k, json.dumps(d[k])[:400])
    except Exception as ex:
        print(f, "unreadable", ex)
PY
