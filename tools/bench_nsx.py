#!/usr/bin/env python
"""Device-resident timing of the fixed-point suppressor kernel (nsx_kernel) at BASELINE config 3's size, every compiled
launch shape, after 600 ticks of ageing; prints one JSON line per shape.  usage: tools/bench_nsx.py [--streams N] [--cfgs 0,1,..]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NSX_BYTES_PER_STREAM_TICK = {16000: 9.93e3, 8000: 5.25e3}   # record read + written (start-up array excluded) + PCM in/out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=100000)
    ap.add_argument("--freq", type=int, default=16000)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--prime", type=int, default=600)
    ap.add_argument("--cfgs", default="0,1,2,3,4,5")
    ap.add_argument("--align", default="1,0", help="nsx_sync values: 1 = the warps of a CTA start every frame together, 0 = free running")
    ap.add_argument("--float-core", action="store_true", help="also time the float core's kernel on the same input")
    ap.add_argument("--float-cfgs", default="-1", help="ns_cfg shapes of the float core to time (-1 = the library's default)")
    a = ap.parse_args()
    import torch

    import wmix_b200
    from wmix_b200.synth import make_frames

    dev = torch.device("cuda", 0)
    peak, peak_src = bench.peaks()
    S, L, R = a.streams, a.freq // 100, 16
    base = make_frames(min(2048, S), a.freq, 300, R, seed=100)
    reps = (S + base.shape[1] - 1) // base.shape[1]
    d_pool = torch.from_numpy(np.ascontiguousarray(np.tile(base, (1, reps, 1))[:, :S])).to(dev)
    d_out = torch.empty((S, L), dtype=torch.int16, device=dev)
    st = torch.cuda.current_stream()
    cores = [1] + ([0] if a.float_core else [])
    for core in cores:
        eng = wmix_b200.Engine(S, a.freq, stages=wmix_b200.NS, ns_core=core)
        for t in range(a.prime):
            eng.tick_device(d_pool[t % R], d_out, None, wmix_b200.NS, st)
        torch.cuda.synchronize()
        variants = [(int(c), int(al)) for c in a.cfgs.split(",") for al in a.align.split(",")] if core == 1 else [(int(c), 1) for c in a.float_cfgs.split(",")]
        for cfg, al in variants:
            if core == 1:
                eng.set_tuning("nsx_cfg", cfg)
                eng.set_tuning("nsx_sync", al)
            elif cfg >= 0:
                eng.set_tuning("ns_cfg", cfg)
            for t in range(10):
                eng.tick_device(d_pool[t % R], d_out, None, wmix_b200.NS, st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for t in range(a.steps):
                eng.tick_device(d_pool[t % R], d_out, None, wmix_b200.NS, st)
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            byt = NSX_BYTES_PER_STREAM_TICK[a.freq] if core == 1 else bench.NS_BYTES_PER_STREAM_TICK
            ach = S * byt / (ms * 1e-3) / 1e9
            print(json.dumps({"core": "nsx" if core else "float", "cfg": cfg, "align": al, "streams": S, "freq": a.freq, "ms_per_tick": round(ms, 4),
                              "achieved_gbs": round(ach, 1), "peak_gbs": peak, "frac": round(ach / peak, 4),
                              "state_bytes_per_stream": eng.state_bytes_per_stream()}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
