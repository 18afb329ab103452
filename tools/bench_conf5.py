#!/usr/bin/env python
"""Measurement of BASELINE config 5 (not the headline bench line): N-minus-one conference mix of G.711
participants, 8 kHz, 8192 participants per GPU (65 536 on 8 GPUs), conferences striped over all ranks so
every bus row crosses NVLink.  Compares the fused peer-memory kernel with bus_sum -> NCCL all-reduce -> nminus1.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_conf5.py [--per-gpu 8192] [--conf-size 1024] [--steps 300] [--warmup 30]

Prints one JSON line on rank 0 (times are CUDA-event times on each rank's stream, max over ranks)."""
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
    ap.add_argument("--conf-size", type=int, default=1024, help="participants per conference (global)")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--law", type=int, default=0)
    ap.add_argument("--rs", type=int, default=-1, help="fused kernel: force the reduce-scatter exchange on (1) / off (0); -1 = the library's choice")
    ap.add_argument("--tile", default="", help="fused kernel: 'row' or 16 (samples per tile); '' = the library's choice")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    from wmix_b200.conference import ConferencePlan, ShardedConference

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = a.per_gpu * world
    n_conf = max(1, total // a.conf_size)
    plan = ConferencePlan([total // n_conf] * n_conf, world, "striped")
    frame = 80
    rng = np.random.default_rng(99 + rank)
    R = 8
    pool = torch.from_numpy(rng.integers(0, 256, (R, plan.local_count(rank), frame)).astype(np.uint8)).to(dev)
    d_out = torch.empty_like(pool[0])
    d_bus = torch.empty((plan.n_conf, frame), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    res = {}
    for mode in ("peer", "nccl", "nccl_c"):
        opts = {}
        if a.rs >= 0:
            opts["reduce_scatter"] = a.rs
        if a.tile:
            opts["tile"] = "row" if a.tile == "row" else int(a.tile)
        conf = ShardedConference(plan, rank, law=a.law, freq=8000, mode=mode, device=local, peer_opts=(opts or None) if mode == "peer" else None)
        for t in range(a.warmup):
            conf.tick(pool[t % R], d_out, d_bus)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for t in range(a.steps):
            conf.tick(pool[t % R], d_out, d_bus)
        e1.record(st)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[mode] = float(ms.item())
        status = conf.status()
        assert status == 0, "peer bus timed out waiting for rank %d" % (status - 1)
        conf.close()
    if rank == 0:
        print(json.dumps({
            "workload": "BASELINE config 5: N-minus-one conference mix, %d G.711 (%s) participants at 8 kHz over %d GPU(s), "
                        "%d conferences of %d striped over all ranks" % (total, "A-law" if a.law == 0 else "mu-law", world, plan.n_conf, total // n_conf),
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "fused_opts": {"reduce_scatter": a.rs, "tile": a.tile or "default"},
            "ms_per_tick": {"peer (one fused kernel, NVLink peer stores)": res["peer"], "nccl (bus_sum -> all_reduce int32 -> nminus1)": res["nccl"],
                            "nccl_c (the same three steps behind wmixb_nccl_bus_tick_device, libnccl opened by the library)": res["nccl_c"]},
            "participants_per_10ms_tick_realtime": {k: total * 10.0 / v for k, v in res.items()},
            "bus_bytes_exchanged_per_rank": plan.n_conf * frame * 4 * (world - 1),
            "algorithmic_bytes_per_participant_tick": 160,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
