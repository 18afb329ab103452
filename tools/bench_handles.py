#!/usr/bin/env python
"""Latency of the drop-in handle calls (one stream per handle, one synchronous round trip per call): microseconds per
ns_process / agc_process / vad_process call on 10 ms packets, after warm-up.  usage: tools/bench_handles.py [--freq 16000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--freq", type=int, default=16000)
    ap.add_argument("--calls", type=int, default=2000)
    a = ap.parse_args()
    import wmix_b200
    from wmix_b200.synth import make_frames

    lib = wmix_b200.lib()
    n = a.freq // 100
    x = np.ascontiguousarray(make_frames(1, a.freq, 300, 64, seed=3)[:, 0])
    out = {}
    for core in (0, 1):
        lib.wmixb_set_default_ns_core(core)
        ns = lib.ns_init(1, a.freq, None)
        agc = lib.agc_init(1, a.freq, 10, 5, None)
        vad = lib.vad_init(1, a.freq, 10, None)
        buf = x[0].copy()
        for name, fn in (("ns", lambda: lib.ns_process(ns, buf.ctypes.data, buf.ctypes.data, n)),
                         ("agc", lambda: lib.agc_process(agc, buf.ctypes.data, buf.ctypes.data, n)),
                         ("vad", lambda: lib.vad_process(vad, buf.ctypes.data, n))):
            for k in range(300):
                buf[:] = x[k % 64]
                fn()
            t0 = time.perf_counter()
            for k in range(a.calls):
                fn()
            out["%s%s_us_per_call" % (name, "x" if core and name == "ns" else "")] = (time.perf_counter() - t0) * 1e6 / a.calls
        lib.ns_release(ns)
        lib.agc_release(agc)
        lib.vad_release(vad)
    lib.wmixb_set_default_ns_core(0)
    print(json.dumps({"freq": a.freq, **{k: round(v, 2) for k, v in out.items()}}))


if __name__ == "__main__":
    main()
