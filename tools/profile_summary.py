#!/usr/bin/env python
"""Write profiles/<tag>_summary.md from what tools/gpu_profile.sh <tag> left in gpurun_out/ (bench lines, launch list,
the three `--set full` captures) and copy the small artefacts next to it.  usage: python tools/profile_summary.py <tag> [notes.md]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) != len(rows[hi]) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        u = r[col["Metric Unit"]]
        v = v / 1000.0 if u in ("nsecond", "ns") else (v * 1000.0 if u in ("msecond", "ms") else v)
        agg[r[col["Kernel Name"]].split("(")[0].replace("void ", "")].append(v)
    tot = sum(sum(v) / len(v) for v in agg.values())
    return [(k, sum(v) / len(v), 100.0 * sum(v) / len(v) / tot) for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]


def main():
    tag = sys.argv[1]
    notes = open(sys.argv[2]).read() if len(sys.argv) > 2 else ""
    g = lambda f: os.path.join(R, "gpurun_out", "%s_%s" % (tag, f))
    b, ref, aec = (json.load(open(g(f))) for f in ("bench.json", "bench_reference.json", "aec.json"))
    summ = lambda k: subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_summary.py"), g(k + ".ncu-rep")],
                                    capture_output=True, text=True).stdout
    km, e, rf = b["kernel_ms"], b["e2e"], b["roofline"]
    out = ["# Round 1, capture %s" % tag.split("_")[-1].upper(), "",
           "Commands (one B200 via `gpurun`, `tools/gpu_profile.sh %s`): `python bench.py` and `python bench.py --impl reference` (not under a" % tag,
           "profiler) -> `%s_bench.json`, `%s_bench_reference.json`; `python tools/bench_aec.py` -> `%s_aec.json`;" % (tag, tag, tag),
           "`ncu --metrics gpu__time_duration.sum --clock-control none -s 750 -c 60 --csv` on the bench command -> `%s_launches.csv`;" % tag,
           "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1` past frame 260 for ns_kernel / post_kernel, past",
           "tick 420 for aec_kernel.", "", notes, "",
           "## Bench lines (CUDA events, not profiled)", "| | |", "|---|---|",
           "| ms per 10 ms tick, 100 000 streams, 1 GPU | %.3f (ns_kernel %.3f, post_kernel %.3f, bus_sum %.3f) |"
           % (b["ms_per_step"], km["ns_kernel"], km["post_kernel(agc+vad)"], km["bus_sum_kernel"]),
           "| real-time streams per GPU (`value`) | %d (headroom %.2fx at 100 k) |" % (b["value"], b["realtime_headroom"]),
           "| roofline, ns_kernel, 14.4 KB / stream-tick | %.0f GB/s = %.3f of the measured %.1f GB/s |" % (rf["achieved"], rf["frac"], rf["peak"]),
           "| e2e, ticks fed back to back (`wmixb_tick_host_submit` / `_wait`; pinned host PCM in -> PCM + VAD flags + bus out) | %.3f ms / tick = %d streams |"
           % (e["ms_per_step"], e["value"]),
           "| e2e, one blocking call per tick (`wmixb_tick_host_bus`) | %.3f ms / tick = %d streams |" % (e["sync_call_ms_per_step"], e["sync_call_value"]),
           "| reference C chain, %d host cores (`--impl reference`) | %d streams |" % (ref["cpu_baseline"]["cores"], ref["value"]),
           "| clocks | %s |" % json.dumps(b["clocks"]),
           "| config 4 (NS -> AEC, 16 384 pairs, 8 kHz) | %.3f ms / tick (aec_kernel %.3f, ns_kernel<128> %.3f); aec_kernel %.0f GB/s = %.3f of peak at 29 KB / stream-tick |"
           % (aec["ms_per_tick"], aec["kernel_ms"]["aec_kernel"], aec["kernel_ms"]["ns_kernel<128>"], aec["roofline"]["achieved"], aec["roofline"]["frac"]),
           "", "## Launch shares (ncu launch list, cold-cache serialised; agrees with the CUDA-event split above)", "| kernel | mean us | share |", "|---|---|---|"]
    out += ["| %s | %.1f | %.1f %% |" % x for x in launch_shares(g("launches.csv"))]
    for k, name in (("ns", "ns"), ("post", "post"), ("aec", "aec")):
        out += ["", "## %s kernel, `--set full`" % name, summ(k)]
    open(os.path.join(R, "profiles", "%s_summary.md" % tag), "w").write("\n".join(out))
    for f in ("bench.json", "bench_reference.json", "aec.json", "launches.csv"):
        shutil.copy(g(f), os.path.join(R, "profiles", "%s_%s" % (tag, f)))


if __name__ == "__main__":
    main()
