#!/usr/bin/env python
"""Small run of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py

NS (16 and 8 kHz), AGC+VAD, AEC, conference bus, G.711, the fused peer-bus kernel (world 1: the tool serialises
launches, so two ranks spinning on each other cannot be run under it), RTP pack/unpack, zoom and mix-load.
Sizes are tiny on purpose: the tools slow kernels down 10-100x."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import wmix_b200
    from wmix_b200 import AEC, AGC, NS, VAD
    from wmix_b200.conference import ConferencePlan, ShardedConference
    from wmix_b200.engine import g711_decode, g711_encode
    from wmix_b200.synth import make_aec_pairs, make_frames

    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream()
    ticks = int(os.environ.get("SAN_TICKS", "3"))
    for freq, S, ns_cfg in ((16000, 83, -1), (8000, 45, -1), (8000, 45, 0)):     # 8 kHz also in the CTA-cooperative shape
        L = freq // 100
        x = make_frames(S, freq, 0, ticks, seed=3)
        eng = wmix_b200.Engine(S, freq)
        if ns_cfg >= 0:
            eng.set_tuning("ns_cfg", ns_cfg)
        eng.set_conferences(np.array([0, 7, 7, 40, S], dtype=np.int32))
        d_in = torch.empty((S, L), dtype=torch.int16, device=dev)
        d_out = torch.empty_like(d_in)
        d_vad = torch.zeros((S,), dtype=torch.uint8, device=dev)
        d_bus = torch.empty((4, L), dtype=torch.int32, device=dev)
        for t in range(ticks):
            d_in.copy_(torch.from_numpy(x[t]))
            eng.tick_device(d_in, d_out, d_vad, 0, st)
            eng.bus_sum(d_out, d_bus, st)
            eng.bus_nminus1(d_bus, d_out, d_in, st)
        # offline mode: several frames per launch
        d_seq = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 0, 2))).to(dev)
        d_seq_out = torch.empty_like(d_seq)
        eng.offline_device(d_seq, d_seq_out, ticks, None, NS | AGC | VAD, st)
        torch.cuda.synchronize()
        eng.close()
        print("ns/agc/vad/bus %d Hz ok" % freq, flush=True)
    if os.environ.get("SAN_ONLY", "") == "ns":
        return
    # the fixed-point suppressor: mono ticks in the default shape and two others (all-zero stream 1 leaves the frame early),
    # offline frames, the second band, a 32 kHz engine
    if os.environ.get("SAN_ONLY", "") in ("", "nsx"):
        for freq, S in ((16000, 83), (8000, 45)):
            L = freq // 100
            x = make_frames(S, freq, 0, ticks, seed=4)
            for cfg, mask in ((7, 1), (0, 1), (4, 0)):
                eng = wmix_b200.Engine(S, freq, stages=NS, ns_core=1)
                eng.set_tuning("nsx_cfg", cfg)
                eng.set_tuning("nsx_sync", mask)
                d_in = torch.empty((S, L), dtype=torch.int16, device=dev)
                for t in range(ticks):
                    d_in.copy_(torch.from_numpy(x[t]))
                    eng.tick_device(d_in, d_in, None, NS, st)
                d_seq = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 0, 2))).to(dev)
                eng.offline_device(d_seq, torch.empty_like(d_seq), ticks, None, NS, st)
                torch.cuda.synchronize()
                eng.close()
            eng = wmix_b200.Engine(S, freq, stages=NS, ns_core=1, ns_high_band=1)
            a = torch.from_numpy(x[0]).to(dev)
            b = torch.from_numpy(x[1]).to(dev)
            for t in range(ticks):
                assert eng.L.wmixb_ns2_device(eng.h, a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), st.cuda_stream) == 0
            torch.cuda.synchronize()
            eng.close()
            print("nsx %d Hz ok" % freq, flush=True)
        eng = wmix_b200.Engine(21, 32000, ns_core=1)
        d = torch.randint(-3000, 3000, (21, 320), dtype=torch.int16, device=dev)
        d_v = torch.zeros((21,), dtype=torch.uint8, device=dev)
        for t in range(ticks):
            eng.tick_device(d, d, d_v, 0, st)
        torch.cuda.synchronize()
        eng.close()
        print("nsx 32 kHz engine ok", flush=True)
        if os.environ.get("SAN_ONLY", "") == "nsx":
            return
    # AEC + NS chain at 8 kHz
    S, L = 37, 80
    far, near = make_aec_pairs(S, 8000, 0, ticks + 2, seed=9)
    eng = wmix_b200.Engine(S, 8000, stages=NS | AEC)
    d_far = torch.empty((S, L), dtype=torch.int16, device=dev)
    d_near = torch.empty_like(d_far)
    d_out = torch.empty_like(d_far)
    for t in range(ticks + 2):
        d_far.copy_(torch.from_numpy(far[t]))
        d_near.copy_(torch.from_numpy(near[t]))
        eng.tick_chain_device(d_far, d_near, d_out, None, NS | AEC, 0, st)
    torch.cuda.synchronize()
    print("aec status", eng.aec_status(), flush=True)
    eng.close()
    # G.711 + fused peer bus (world 1), both exchange shapes
    for sizes, opts in (([5, 0, 17, 3, 1], None), ([4] * 300, {"tile": "row"})):
        plan = ConferencePlan(sizes, 1)
        n = plan.local_count(0)
        conf = ShardedConference(plan, 0, law=0, freq=8000, mode="peer", device=0, peer_opts=opts)
        codes = torch.randint(0, 256, (n, 80), dtype=torch.uint8, device=dev)
        d_o = torch.empty_like(codes)
        d_b = torch.empty((plan.n_conf, 80), dtype=torch.int32, device=dev)
        for t in range(ticks):
            conf.tick(codes, d_o, d_b)
        torch.cuda.synchronize()
        assert conf.status() == 0
        conf.close()
    pcm = torch.randint(-32768, 32767, (4096,), dtype=torch.int16, device=dev)
    c = torch.empty((4096,), dtype=torch.uint8, device=dev)
    g711_encode(0, pcm, c, 4096, st)
    g711_decode(1, c, pcm, 4096, st)
    torch.cuda.synchronize()
    print("g711 / peer bus ok", flush=True)
    # drop-in handle API, one stream
    lib = wmix_b200.lib()
    h = lib.ns_init(1, 16000, None)
    buf = np.ascontiguousarray(make_frames(1, 16000, 0, 2, seed=1)[:, 0].reshape(-1))
    lib.ns_process(h, buf.ctypes.data, buf.ctypes.data, 320)
    lib.ns_release(h)
    print("handles ok", flush=True)


if __name__ == "__main__":
    main()
