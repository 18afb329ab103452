#!/usr/bin/env python
"""One GPU, no peers: the fused conference-bus kernel (world = 1) against bus_sum -> nminus1 on the same legs — what the
fused kernel costs before any exchange.  usage: python tools/bench_conf5_local.py [--per-gpu 8192] [--steps 300]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    a = ap.parse_args()
    import torch

    from wmix_b200.conference import ConferencePlan, ShardedConference

    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream()
    rng = np.random.default_rng(5)
    res = {}
    for conf_local in (512, 8, 2):                       # local members per conference (1024 / 16 / 16 striped over 2 and 8 ranks)
        n_conf = a.per_gpu // conf_local
        plan = ConferencePlan([conf_local] * n_conf, 1)
        pool = torch.from_numpy(rng.integers(0, 256, (8, a.per_gpu, 80)).astype(np.uint8)).to(dev)
        d_out = torch.empty_like(pool[0])
        d_bus = torch.empty((n_conf, 80), dtype=torch.int32, device=dev)
        for mode in ("peer", "local"):
            conf = ShardedConference(plan, 0, law=0, freq=8000, mode=mode, device=0)
            for t in range(a.warmup):
                conf.tick(pool[t % 8], d_out, d_bus)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for t in range(a.steps):
                conf.tick(pool[t % 8], d_out, d_bus)
            e1.record(st)
            torch.cuda.synchronize()
            res["%d conferences x %d local members, %s" % (n_conf, conf_local, mode)] = e0.elapsed_time(e1) / a.steps * 1e3
            conf.close()
    print(json.dumps({"us_per_tick": res}))


if __name__ == "__main__":
    main()
