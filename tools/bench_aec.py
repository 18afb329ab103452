#!/usr/bin/env python
"""Measurement of BASELINE config 4 (not the headline bench line): NS -> AEC on N near/far pairs, 8 kHz mono,
10 ms ticks, device-resident PCM.  Prints one JSON line: ms per tick, real-time streams, and the AEC kernel's
achieved algorithmic GB/s (~29 KB per stream-tick) against the measured HBM peak.

    python tools/bench_aec.py [--streams 16384] [--steps 300] [--warmup 400] [--ring 8]

The warm-up runs past the AEC's start-up phase (buffer-size settling) so the timed ticks are the NLMS steady state."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

AEC_BYTES = 29.0e3      # ~23 KB per 64-sample block x 1.25 blocks per 10 ms tick
NS8_BYTES = 7.1e3       # NS at 8 kHz incl. PCM


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=400)
    ap.add_argument("--ring", type=int, default=8)
    ap.add_argument("--no-ns", action="store_true")
    ap.add_argument("--aec-align", type=int, default=-1, help="experiment: CTA alignment of the AEC kernel (wmixb_set_tuning)")
    ap.add_argument("--aec-warps", type=int, default=-1, help="experiment: warps per CTA of the AEC kernel, 8 or 16 (wmixb_set_tuning)")
    a = ap.parse_args()
    import torch

    import wmix_b200
    from wmix_b200 import AEC, NS
    from wmix_b200.synth import make_aec_pairs

    dev = torch.device("cuda", 0)
    S, L = a.streams, 80
    base = min(1024, S)
    T = a.warmup + a.steps
    # distinct seeded pairs for `base` streams, tiled; consecutive ticks so the echo path is coherent in time
    far, near = make_aec_pairs(base, 8000, 0, T, seed=41)
    reps = (S + base - 1) // base
    stages = AEC | (0 if a.no_ns else NS)
    eng = wmix_b200.Engine(S, 8000, stages=stages)
    if a.aec_align >= 0:
        eng.set_tuning("aec_align", a.aec_align)
    if a.aec_warps > 0:
        eng.set_tuning("aec_warps", a.aec_warps)
    d_far = torch.empty((S, L), dtype=torch.int16, device=dev)
    d_near = torch.empty((S, L), dtype=torch.int16, device=dev)
    d_out = torch.empty((S, L), dtype=torch.int16, device=dev)
    far_d = torch.from_numpy(far).to(dev)
    near_d = torch.from_numpy(near).to(dev)
    st = torch.cuda.current_stream()

    def load(t):
        d_far.copy_(far_d[t].repeat(reps, 1)[:S])
        d_near.copy_(near_d[t].repeat(reps, 1)[:S])

    for t in range(a.warmup):
        load(t)
        if a.no_ns:
            eng.aec_device(d_far, d_near, d_out, L, 0, st)
        else:
            eng.tick_chain_device(d_far, d_near, d_out, None, NS | AEC, 0, st)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    for k in range(a.steps):
        load(a.warmup + k)
        ev[k][0].record(st)
        if not a.no_ns:
            eng.tick_device(d_near, d_out, None, NS, st)
        ev[k][1].record(st)
        eng.aec_device(d_far, d_near if a.no_ns else d_out, d_out, L, 0, st)
        ev[k][2].record(st)
    torch.cuda.synchronize()
    ns_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / a.steps
    aec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / a.steps
    flags = eng.aec_status()
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    ach = S * AEC_BYTES / (aec_ms * 1e-3) / 1e9
    print(json.dumps({
        "workload": "BASELINE config 4: %sAEC (PBFDAF NLMS), 8 kHz mono near/far pairs, %d streams" % ("" if a.no_ns else "NS -> ", S),
        "steps": a.steps, "warmup": a.warmup, "ms_per_tick": ns_ms + aec_ms,
        "kernel_ms": {"ns_kernel<128>": ns_ms, "aec_kernel": aec_ms},
        "realtime_streams": S * 10.0 / (ns_ms + aec_ms), "realtime_headroom": 10.0 / (ns_ms + aec_ms),
        "roofline": {"bound": "hbm", "kernel": "aec_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "algorithmic_bytes_per_launch": S * AEC_BYTES},
        "aec_status": {"flags": flags[0], "flagged_streams": flags[1]},
        "state_bytes_per_stream": eng.state_bytes_per_stream(),
    }))
    eng.close()


if __name__ == "__main__":
    main()
