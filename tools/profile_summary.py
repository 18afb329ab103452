#!/usr/bin/env python
"""Write profiles/<tag>_summary.md from what tools/gpu_profile.sh <tag> left in gpurun_out/ (bench lines, launch list,
the `--set full` captures) and copy the small artefacts next to it.  usage: python tools/profile_summary.py <tag> [notes.md]"""
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


def last_json(path):
    return json.loads([ln for ln in open(path) if ln.startswith("{")][-1])


def main():
    tag = sys.argv[1]
    notes = open(sys.argv[2]).read() if len(sys.argv) > 2 else ""
    g = lambda f: os.path.join(R, "gpurun_out", "%s_%s" % (tag, f))
    b, ref = last_json(g("bench.json")), last_json(g("bench_reference.json"))
    w5 = last_json(g("bench_w5.json")) if os.path.exists(g("bench_w5.json")) else None
    def summ(k):
        # the summary written on the GPU box (tools/gpu_profile.sh) when the report itself did not travel back
        if os.path.exists(g(k + "_ncu.md")):
            return open(g(k + "_ncu.md")).read()
        return subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_summary.py"), g(k + ".ncu-rep")], capture_output=True, text=True).stdout
    km, e, rf = b["kernel_ms"], b["e2e"], b["roofline"]
    out = ["# Round 2, capture %s" % tag.split("_")[-1].upper(), "",
           "Commands (one B200 via `gpurun`, `tools/gpu_profile.sh %s`): `python bench.py`, `python bench.py --steps 20 --warmup 5` (the driver's flags) and" % tag,
           "`python bench.py --impl reference --steps 20 --warmup 5`, not under a profiler -> `%s_bench.json`, `%s_bench_w5.json`, `%s_bench_reference.json`;" % (tag, tag, tag),
           "`ncu --metrics gpu__time_duration.sum --clock-control none -s 1840 -c 60 --csv` on the bench command -> `%s_launches.csv`;" % tag,
           "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1` past tick 605 (both arms age every handle 600 ticks) for",
           "ns_cta_kernel / post_kernel, past tick 420 for aec_kernel.", "", notes, "",
           "## Bench lines (CUDA events, not profiled)", "| | |", "|---|---|",
           "| ms per 10 ms tick, 100 000 streams, 1 GPU | %.3f (ns %.3f, post %.3f, bus_sum %.3f) |"
           % (b["ms_per_step"], km["ns_kernel"], km["post_kernel(agc+vad)"], km["bus_sum_kernel"]),
           "| real-time streams per GPU (`value`) | %d (headroom %.2fx at 100 k) |" % (b["value"], b["realtime_headroom"]),
           "| roofline, NS kernel, 14.4 KB / stream-tick | %.0f GB/s = %.3f of the measured %.1f GB/s |" % (rf["achieved"], rf["frac"], rf["peak"]),
           "| e2e, ticks fed back to back (pinned host PCM in -> PCM + VAD flags + bus out) | %.3f ms / tick = %d streams; bare copies of the same bytes: %.3f ms (e2e at %.2f of that ceiling) |"
           % (e["ms_per_step"], e["value"], e["copy_ceiling_ms_per_step"], e["frac_of_copy_ceiling"]),
           "| e2e, bus + VAD flags only come back (`h_out = NULL`) | %.3f ms / tick = %d streams |" % (e["bus_only"]["ms_per_step"], e["bus_only"]["value"]),
           "| e2e, one blocking call per tick (`wmixb_tick_host_bus`) | %.3f ms / tick = %d streams |" % (e["sync_call_ms_per_step"], e["sync_call_value"]),
           "| reference C chain, %d host cores (`--impl reference`, same 600-tick ageing) | %d streams |" % (ref["cpu_baseline"]["cores"], ref["value"]),
           "| clocks | %s |" % json.dumps(b["clocks"])]
    if w5:
        out.append("| the same line with the driver's flags (`--steps 20 --warmup 5`) | value %d, ns %.3f ms, e2e %d |"
                   % (w5["value"], w5["kernel_ms"]["ns_kernel"], w5["e2e"]["value"]))
    if "full_load" in b:
        fl = b["full_load"]
        out.append("| %d streams resident on the GPU (state %.1f GB) | %.2f ms per tick (fits the 10 ms tick: %s) |"
                   % (fl["streams_resident"], fl["state_gb"], fl["ms_per_tick"], fl["fits_10ms_tick"]))
    if "config4" in b:
        c4 = b["config4"]
        out.append("| config 4 (NS -> AEC, 16 384 pairs, 8 kHz) | %.3f ms / tick (aec_kernel %.3f, ns_kernel<128> %.3f); aec_kernel %.0f GB/s = %.3f of peak at 29 KB / stream-tick |"
                   % (c4["ms_per_tick"], c4["kernel_ms"]["aec_kernel"], c4["kernel_ms"]["ns_kernel<128>"], c4["roofline"]["achieved"], c4["roofline"]["frac"]))
    if "offline" in b:
        o = b["offline"]
        out.append("| persistent offline mode, %d streams x %d frames, NS->AGC->VAD | %.1f M stream-frames/s against %.1f M as ticks (%.2fx); %.2fx of the same call with the records left in HBM |"
                   % (o["streams"], o["frames_per_stream"], o["stream_frames_per_s"]["offline"] / 1e6, o["stream_frames_per_s"]["ticks"] / 1e6,
                      o["offline_vs_ticks"], o["staged_vs_unstaged"]))
    out += ["", "## Launch shares (ncu launch list, cold-cache serialised; agrees with the CUDA-event split above)", "| kernel | mean us | share |", "|---|---|---|"]
    out += ["| %s | %.1f | %.1f %% |" % x for x in launch_shares(g("launches.csv"))]
    for k in ("ns", "post", "aec"):
        if os.path.exists(g(k + ".ncu-rep")) or os.path.exists(g(k + "_ncu.md")):
            out += ["", "## %s kernel, `--set full`" % k, summ(k)]
    open(os.path.join(R, "profiles", "%s_summary.md" % tag), "w").write("\n".join(out))
    for f in ("bench.json", "bench_w5.json", "bench_reference.json", "launches.csv", "tests.txt"):
        if os.path.exists(g(f)):
            shutil.copy(g(f), os.path.join(R, "profiles", "%s_%s" % (tag, f)))


if __name__ == "__main__":
    main()
