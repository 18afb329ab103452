#!/usr/bin/env python
"""bench.py — headline benchmark of wmix_b200 (contract: see the task prompt / DESIGN.md §6).

Metric (BASELINE.json): real-time 16 kHz mono streams sustained per GPU through
NS -> AGC -> VAD -> conference-bus mix inside a 10 ms tick, and the fraction of the measured
B200 HBM roofline the dominant kernel (ns_kernel) reaches.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA)
    python bench.py --impl reference --gpus N --steps K ...   # the reference C path on host cores

A "step" is one 10 ms tick over every stream of the job: ns_kernel, post_kernel (AGC+VAD) and
bus_sum_kernel.  Workload: BASELINE config 3, 100 000 streams per GPU (weak scaling: every rank
owns its own 100 000 streams and its own conferences; that path has no cross-stream exchange).
value = streams_total * 10 ms / ms_per_step.

Both arms age every handle by the same PRIME = 600 untimed ticks before the W warm-up ticks and the K
timed ones ("timed after 600 warm-up"), whatever --warmup says: WebRtcNs is
1.5-2x slower per frame inside its 50-frame start-up model and switches its gain map on at frame 200, so a
run timed on fresh handles measures a different regime on both sides.

Keyed sub-measurements in the same JSON line (none of them changes `value`):
  conf5     (N > 1)  BASELINE config 5, the one collective of the path: the conference bus striped over all
                     ranks — fused peer-memory kernel vs bus_sum -> NCCL all-reduce -> nminus1, with an in-run
                     check of the exchanged bus against a torch int32 sum of all ranks' legs (parity_ok);
  config4   (N = 1)  BASELINE config 4: NS -> AEC on 16 384 near/far pairs at 8 kHz;
  full_load          >= 1 000 000 resident streams per GPU through one tick, so that `value` is backed by a
                     run at that stream count and not only by extrapolation from 100 000;
  nsx       (N = 1)  the same tick with the reference's other suppressor, the fixed-point WebRtcNsx_* core (ns_core = 1);
  offline   (N = 1)  the persistent offline mode (K frames per stream per launch, NS state on chip) against K ticks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FREQ = 16000
FRAME = 160
CONF_SIZE = 16
PRIME = 600                            # untimed ticks on every handle of BOTH arms before warm-up (config 3)
METRIC = "real-time 16 kHz streams per GPU, NS+VAD+AGC+mix, 10 ms tick; % HBM roofline"   # BASELINE.json, both arms
UNIT = "real-time 16 kHz streams (10 ms tick), whole job"                                  # both arms
NS_BYTES_PER_STREAM_TICK = 14.4e3      # NS state R+W + PCM in/out
CHAIN_BYTES_PER_STREAM_TICK = 16.0e3   # NS + VAD + AGC + mix
AEC_BYTES_PER_STREAM_TICK = 29.0e3     # AEC at 8 kHz
NSX_BYTES_PER_STREAM_TICK = 9.93e3     # fixed-point NS: 4.64 KB of record read + 4.64 KB written + PCM in/out
# dram__bytes_read.sum + dram__bytes_write.sum of one NS launch per stream: a CONSTANT taken from the latest
# `ncu --set full` capture (it cannot be measured inside an unprofiled run); see NS_TRAFFIC_SOURCE
NS_DRAM_TRAFFIC_PER_STREAM_NCU = (874.530048e6 + 650.368000e6) / 100_000
NS_TRAFFIC_SOURCE = "constant from ncu --set full capture r2zz (profiles/r2zz_summary.md), not measured in this run"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_pool(n_streams, n_ring, seed):
    """[n_ring, n_streams, FRAME] int16: n_ring consecutive ticks (speech active) of 2048 distinct seeded
    streams, tiled over n_streams.  Synthetic speech + noise, SURVEY.md §8(d)."""
    from wmix_b200.synth import make_frames

    base = min(2048, n_streams)
    x = make_frames(base, FREQ, 300, n_ring, seed=seed)                   # [R, base, L]
    reps = (n_streams + base - 1) // base
    return np.ascontiguousarray(np.tile(x, (1, reps, 1))[:, :n_streams])


def cpu_leg(n_streams, n_ticks, kind_pref="reference", prime=PRIME):
    """The reference C chain on the host cores (bounded sample), timed after `prime` untimed ticks on the same handles —
    the regime the GPU arm is timed in (past the suppressor's start-up model and gain-map switch).  Returns dict for
    cpu_baseline."""
    from tests._oracle import oracle, P

    L = oracle()
    L.orc_bench_chain_primed.restype = C.c_double
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libwmix_ref.so")
    kind = "reference" if (kind_pref == "reference" and os.path.exists(ref_so)) else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_streams = max(CONF_SIZE * cores, n_streams // (CONF_SIZE * cores) * (CONF_SIZE * cores))
    x = make_pool(n_streams, n_ticks, seed=7)
    bus = np.zeros((n_ticks, n_streams // CONF_SIZE, FRAME), np.int32)
    sec = L.orc_bench_chain_primed(ref_so.encode() if kind == "reference" else None, FREQ, n_streams, int(prime), n_ticks, CONF_SIZE,
                                   cores, P(x), None, P(bus))
    if sec <= 0:
        raise RuntimeError("orc_bench_chain failed: %r" % sec)
    ms_per_tick = sec * 1e3 / n_ticks
    return {"value": n_streams * 10.0 / ms_per_tick, "unit": UNIT, "cores": cores,
            "kind": kind, "sample": "%d streams x %d ticks (after %d untimed ticks on the same handles), NS->AGC->VAD->bus, -O2 build, one pthread per core"
                      % (n_streams, n_ticks, prime),
            "ms_per_tick": ms_per_tick, "us_per_stream_tick_per_core": sec * 1e6 * cores / (n_streams * n_ticks)}


def workload_text(streams_per_gpu):
    return ("BASELINE config 3: NS->AGC(5 dB)->VAD(mode 3)->int32 conference bus, 16 kHz mono, %d streams per GPU in conferences "
            "of %d; every handle aged %d untimed ticks before warm-up" % (streams_per_gpu, CONF_SIZE, PRIME))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_streams = CONF_SIZE * cores * max(1, 2048 // (CONF_SIZE * cores))
    # one "step" = one tick over the bounded sample; PRIME + W untimed ticks run first on the same handles
    steps = min(args.steps, 200)
    warm = min(max(args.warmup, 0), 300)
    leg = cpu_leg(n_streams, steps, prime=PRIME + warm)
    line = {"impl": "reference", "metric": METRIC,
            "value": leg["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": leg["ms_per_tick"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i16 (float NS with double transcendentals, integer AGC/VAD/mix)", "data": "synthetic",
            "config": {"workload": workload_text(args.streams),
                       "sample": "CPU arm runs a bounded sample of that workload: %s" % leg["sample"],
                       "note": "the unmodified reference (oracle/_ref) when it was built, else the C port; value = streams this "
                               "host sustains in real time"},
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# sub-measurements
# ---------------------------------------------------------------------------------------------------------------
def conf5_leg(torch, dist, world, rank, local, dev, conf_size, steps=200, warmup=30, per_gpu=8192, law=0):
    """BASELINE config 5 on all ranks: `per_gpu` A-law legs per GPU at 8 kHz, conferences of `conf_size` striped over the
    ranks so that every bus row crosses NVLink.  Returns {peer_us, nccl_us, bus_bytes, parity_ok, ...} (rank-max times)."""
    from wmix_b200.conference import ConferencePlan, ShardedConference
    from wmix_b200.engine import g711_decode, g711_encode

    frame = 80
    total = per_gpu * world
    n_conf = max(1, total // conf_size)
    per_conf_local = total // n_conf // world
    plan = ConferencePlan([total // n_conf] * n_conf, world, "striped")
    n_local = plan.local_count(rank)
    g = torch.Generator(device="cpu").manual_seed(991 + rank)
    R = 4
    pool = torch.randint(0, 256, (R, n_local, frame), generator=g, dtype=torch.uint8).to(dev)
    st = torch.cuda.current_stream()
    res, outs, buses = {}, {}, {}
    for mode in ("peer", "nccl", "nccl_c"):
        conf = ShardedConference(plan, rank, law=law, freq=8000, mode=mode, device=local)
        d_out = torch.empty_like(pool[0])
        d_bus = torch.empty((plan.n_conf, frame), dtype=torch.int32, device=dev)
        for t in range(warmup):
            conf.tick(pool[t % R], d_out, d_bus)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for t in range(steps):
            conf.tick(pool[t % R], d_out, d_bus)
        e1.record(st)
        dist.barrier()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / steps * 1e3
        # one more tick on a known input for the parity check
        conf.tick(pool[0], d_out, d_bus)
        torch.cuda.synchronize()
        outs[mode], buses[mode] = d_out.clone(), d_bus.clone()
        healthy = conf.status() == 0
        dist.barrier()
        conf.close()
        res[mode + "_ok"] = healthy
    # in-run parity: the exchanged bus against an int32 sum, computed here with torch, of the decoded legs of ALL ranks
    pcm = torch.empty((n_local, frame), dtype=torch.int16, device=dev)
    g711_decode(law, pool[0], pcm, n_local * frame, st)
    torch.cuda.synchronize()
    # (NCCL has no int16: the decoded legs travel as bytes)
    gathered_b = [torch.empty((n_local, frame * 2), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered_b, pcm.view(torch.uint8))
    gathered = [g_.view(torch.int16) for g_ in gathered_b]
    # striped plan with equal sizes: rank r hosts per_conf_local members of every conference, conference-major
    want_bus = torch.zeros((n_conf, frame), dtype=torch.int32, device=dev)
    for r in range(world):
        want_bus += gathered[r].view(n_conf, per_conf_local, frame).to(torch.int32).sum(dim=1)
    own = pcm.view(n_conf, per_conf_local, frame).to(torch.int32)
    want_pcm = (want_bus[:, None, :] - own).clamp_(-32768, 32767).to(torch.int16).reshape(n_local, frame).contiguous()
    want_codes = torch.empty((n_local, frame), dtype=torch.uint8, device=dev)
    g711_encode(law, want_pcm, want_codes, n_local * frame, st)
    torch.cuda.synchronize()
    ok = all(bool(torch.equal(buses[m], want_bus)) and bool(torch.equal(outs[m], want_codes)) and res[m + "_ok"]
             for m in ("peer", "nccl", "nccl_c"))
    stats = torch.tensor([res["peer"], res["nccl"], res["nccl_c"], 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    peer_us, nccl_us, nccl_c_us, bad = [float(v) for v in stats.cpu()]
    return {"participants": total, "conferences": n_conf, "conference_size": total // n_conf, "law": "A-law" if law == 0 else "mu-law",
            "peer_us": peer_us, "nccl_us": nccl_us, "nccl_c_us": nccl_c_us, "fused_le_nccl": peer_us <= min(nccl_us, nccl_c_us),
            "bus_bytes": n_conf * frame * 4, "nvlink_bytes_per_rank_per_tick": n_conf * frame * 4 * (world - 1),
            "parity_ok": bad == 0.0, "steps": steps,
            "participants_realtime_per_10ms_tick": {"peer": total * 10.0 / (peer_us * 1e-3), "nccl": total * 10.0 / (nccl_us * 1e-3),
                                                    "nccl_c": total * 10.0 / (nccl_c_us * 1e-3)},
            "paths": {"peer": "wmixb_peer_bus_tick_device: one fused kernel per rank, partial rows stored into every peer's mailbox over NVLink",
                      "nccl": "wmixb_g711_bus_sum_device -> torch.distributed all_reduce(int32, SUM) -> wmixb_g711_nminus1_device",
                      "nccl_c": "wmixb_nccl_bus_tick_device: the same three steps on ONE stream behind the C-ABI, ncclAllReduce from the "
                                "library's own communicator (libnccl opened at run time)"}}


def config4_leg(torch, dev, local, peak, steps=60, warmup=420, streams=16384):
    """BASELINE config 4: NS -> AEC on near/far pairs at 8 kHz, timed past the AEC's start-up phase."""
    import wmix_b200
    from wmix_b200 import AEC, NS
    from wmix_b200.synth import make_aec_pairs

    L = 80
    base = 256
    T = warmup + steps
    far, near = make_aec_pairs(base, 8000, 0, T, seed=41)
    reps = (streams + base - 1) // base
    eng = wmix_b200.Engine(streams, 8000, stages=AEC | NS, device=local)
    d_far = torch.empty((streams, L), dtype=torch.int16, device=dev)
    d_near = torch.empty((streams, L), dtype=torch.int16, device=dev)
    d_out = torch.empty((streams, L), dtype=torch.int16, device=dev)
    far_d, near_d = torch.from_numpy(far).to(dev), torch.from_numpy(near).to(dev)
    st = torch.cuda.current_stream()

    def load(t):
        d_far.copy_(far_d[t].repeat(reps, 1)[:streams])
        d_near.copy_(near_d[t].repeat(reps, 1)[:streams])

    for t in range(warmup):
        load(t)
        eng.tick_chain_device(d_far, d_near, d_out, None, NS | AEC, 0, st)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        load(warmup + k)
        ev[k][0].record(st)
        eng.tick_device(d_near, d_out, None, NS, st)
        ev[k][1].record(st)
        eng.aec_device(d_far, d_out, d_out, L, 0, st)
        ev[k][2].record(st)
    torch.cuda.synchronize()
    ns_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    aec_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    flags = eng.aec_status()
    eng.close()
    ach = streams * AEC_BYTES_PER_STREAM_TICK / (aec_ms * 1e-3) / 1e9
    return {"workload": "BASELINE config 4: NS -> AEC (PBFDAF NLMS), 8 kHz mono near/far pairs, %d streams, timed after %d ticks" % (streams, warmup),
            "ms_per_tick": ns_ms + aec_ms, "kernel_ms": {"ns_kernel<128>": ns_ms, "aec_kernel": aec_ms},
            "realtime_streams": streams * 10.0 / (ns_ms + aec_ms), "steps": steps,
            "roofline": {"bound": "hbm", "kernel": "aec_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_launch": streams * AEC_BYTES_PER_STREAM_TICK},
            "aec_status": {"flags": flags[0], "flagged_streams": flags[1]}}


def nsx_leg(torch, dev, local, peak, streams, steps=60):
    """The same tick with the reference's OTHER suppressor, the fixed-point WebRtcNsx_* core (its `#define MAKE_WEBRTC_NSX`,
    R:src/webrtc.c:511-523; wmixb_config.ns_core = 1): NSX -> AGC -> VAD -> bus, aged PRIME ticks, CUDA events per kernel."""
    import wmix_b200

    NS, AGC, VAD = wmix_b200.NS, wmix_b200.AGC, wmix_b200.VAD
    eng = wmix_b200.Engine(streams, FREQ, device=local, ns_core=1)
    eng.set_conferences(np.arange(0, streams + 1, CONF_SIZE, dtype=np.int32))
    R = 8
    base = make_pool(min(2048, streams), R, seed=700)
    reps = (streams + base.shape[1] - 1) // base.shape[1]
    d_pool = torch.from_numpy(base).to(dev).repeat(1, reps, 1)[:, :streams].contiguous()
    d_pcm = torch.empty((streams, FRAME), dtype=torch.int16, device=dev)
    d_vad = torch.zeros((streams,), dtype=torch.uint8, device=dev)
    d_bus = torch.empty((streams // CONF_SIZE, FRAME), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def step(t, ev=None):
        if ev:
            ev[0].record(st)
        eng.tick_device(d_pool[t % R], d_pcm, None, NS, st)
        if ev:
            ev[1].record(st)
        eng.tick_device(d_pcm, d_pcm, d_vad, AGC | VAD, st)
        eng.bus_sum(d_pcm, d_bus, st)
        if ev:
            ev[2].record(st)

    for t in range(PRIME):
        step(t)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        step(PRIME + k, evs[k])
    torch.cuda.synchronize()
    nsx_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    rest_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    state = eng.state_bytes_per_stream()
    eng.close()
    ach = streams * NSX_BYTES_PER_STREAM_TICK / (nsx_ms * 1e-3) / 1e9
    return {"workload": "config 3 with the fixed-point suppressor: NSX->AGC->VAD->bus, 16 kHz mono, %d streams, aged %d ticks" % (streams, PRIME),
            "ms_per_tick": nsx_ms + rest_ms, "realtime_streams": streams * 10.0 / (nsx_ms + rest_ms), "steps": steps,
            "kernel_ms": {"nsx_kernel<256, 32, 1>": nsx_ms, "post_kernel + bus_sum_kernel": rest_ms},
            "state_bytes_per_stream": state,
            "roofline": {"bound": "hbm", "kernel": "nsx_kernel<256, 32, 1>", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "algorithmic_bytes_per_launch": streams * NSX_BYTES_PER_STREAM_TICK,
                         "note": "integer pipeline, bit-exact; bound by instruction issue (~5 500 warp instructions per stream-frame, "
                                 "two 256-point complex 16-bit FFTs), not by HBM — profiles/r2_p_nsx_summary.md"}}


def full_load_leg(torch, dev, local, streams, steps=10):
    """`streams` resident streams on this GPU through the whole tick (NS, AGC+VAD, bus), aged PRIME ticks: is one tick
    inside the 10 ms budget at that stream count?"""
    import wmix_b200

    NS, AGC, VAD = wmix_b200.NS, wmix_b200.AGC, wmix_b200.VAD
    eng = wmix_b200.Engine(streams, FREQ, device=local)
    eng.set_conferences(np.arange(0, streams + 1, CONF_SIZE, dtype=np.int32))
    R = 2
    base = make_pool(2048, R, seed=300)
    reps = (streams + 2047) // 2048
    d_pool = torch.from_numpy(base).to(dev).repeat(1, reps, 1)[:, :streams].contiguous()
    d_pcm = torch.empty((streams, FRAME), dtype=torch.int16, device=dev)
    d_vad = torch.zeros((streams,), dtype=torch.uint8, device=dev)
    d_bus = torch.empty((streams // CONF_SIZE, FRAME), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def step(t):
        eng.tick_device(d_pool[t % R], d_pcm, None, NS, st)
        eng.tick_device(d_pcm, d_pcm, d_vad, AGC | VAD, st)
        eng.bus_sum(d_pcm, d_bus, st)

    for t in range(PRIME):
        step(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for t in range(steps):
        step(PRIME + t)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    state_gb = streams * eng.state_bytes_per_stream() / 1e9
    eng.close()
    del d_pool, d_pcm, d_vad, d_bus
    torch.cuda.empty_cache()
    return {"streams_resident": streams, "ms_per_tick": ms, "fits_10ms_tick": ms <= 10.0, "state_gb": state_gb, "steps": steps,
            "note": "measured at this stream count on this GPU (not extrapolated); PCM resident in HBM"}


def offline_leg(torch, dev, local, streams=41440, K=60):
    """Persistent offline mode against tick mode on the same frames: `streams` streams x K consecutive frames through
    NS -> AGC -> VAD, (a) as K ticks (three launches each; every frame reads and writes its 14 KB record in HBM) and (b) as ONE
    wmixb_offline_device call (the NS record of a stream is pulled into shared memory by one TMA bulk copy, all K frames run
    against it, one bulk store writes it back).  Both after PRIME ticks of ageing; stream-frames per second, CUDA events."""
    import wmix_b200

    NS, AGC, VAD = wmix_b200.NS, wmix_b200.AGC, wmix_b200.VAD
    base = make_pool(2048, K, seed=500)                                            # [K, 2048, L]
    reps = (streams + 2047) // 2048
    d_ticks = torch.from_numpy(base).to(dev).repeat(1, reps, 1)[:, :streams].contiguous()          # [K][S][L]
    d_seq = d_ticks.permute(1, 0, 2).contiguous()                                                   # [S][K][L]
    d_out = torch.empty((streams, FRAME), dtype=torch.int16, device=dev)
    d_seq_out = torch.empty_like(d_seq)
    d_vad = torch.zeros((streams,), dtype=torch.uint8, device=dev)
    d_vad_seq = torch.zeros((streams, K), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    res = {}
    for mode in ("ticks", "offline", "offline_unstaged"):
        eng = wmix_b200.Engine(streams, FREQ, device=local)
        if mode == "offline_unstaged":
            eng.set_tuning("ns_offline_staged", 0)
        for t in range(PRIME):
            eng.tick_device(d_ticks[t % K], d_out, d_vad, NS | AGC | VAD, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps_t = 3
        e0.record(st)
        for _ in range(reps_t):
            if mode == "ticks":
                for f in range(K):
                    eng.tick_device(d_ticks[f], d_out, d_vad, NS | AGC | VAD, st)
            else:
                eng.offline_device(d_seq, d_seq_out, K, d_vad_seq, NS | AGC | VAD, st)
        e1.record(st)
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / reps_t
        eng.close()
    frames = streams * K
    return {"streams": streams, "frames_per_stream": K, "stages": "NS->AGC->VAD",
            "ms": res, "stream_frames_per_s": {k: frames / (v * 1e-3) for k, v in res.items()},
            "offline_vs_ticks": res["ticks"] / res["offline"], "staged_vs_unstaged": res["offline_unstaged"] / res["offline"],
            "note": "offline = wmixb_offline_device, NS records staged in shared memory by TMA bulk copies for the whole run of frames; "
                    "offline_unstaged = the same call with every frame going back to the record in HBM"}


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import wmix_b200
    from wmix_b200.engine import HostBuffer, host_copy_ceiling

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S = args.streams
    n_conf = S // CONF_SIZE
    eng = wmix_b200.Engine(S, FREQ, device=local)
    eng.set_conferences(np.arange(0, S + 1, CONF_SIZE, dtype=np.int32))
    if args.ns_cfg >= 0:
        eng.set_tuning("ns_cfg", args.ns_cfg)            # kernel-shape experiments (tools/); the default is the library's
    if args.post_occ >= 0:
        eng.set_tuning("post_occ", args.post_occ)
    for kv in args.tune:
        k, v = kv.split("=")
        eng.set_tuning(k, int(v))
    R = args.ring
    # tick inputs live in pinned host memory placed for this GPU (wmixb_host_alloc), the device copy is made from it
    h_pool = HostBuffer((R, S, FRAME), np.int16, device=local)
    h_pool.array[...] = make_pool(S, R, seed=100 + rank)
    d_pool = torch.from_numpy(h_pool.array).to(dev)
    d_pcm = torch.empty((S, FRAME), dtype=torch.int16, device=dev)
    d_vad = torch.zeros((S,), dtype=torch.uint8, device=dev)
    d_bus = torch.empty((n_conf, FRAME), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    NS, AGC, VAD = wmix_b200.NS, wmix_b200.AGC, wmix_b200.VAD

    def step(t, evs=None):
        src = d_pool[t % R]
        if evs is not None:
            evs[0].record(stream)
        eng.tick_device(src, d_pcm, None, NS, stream)
        if evs is not None:
            evs[1].record(stream)
        eng.tick_device(d_pcm, d_pcm, d_vad, AGC | VAD, stream)
        if evs is not None:
            evs[2].record(stream)
        eng.bus_sum(d_pcm, d_bus, stream)
        if evs is not None:
            evs[3].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(PRIME):               # ageing, the same for every --warmup
        step(t)
    for t in range(args.warmup):
        step(PRIME + t)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = wmix_b200.kernel_launches()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for k in range(args.steps):
        step(PRIME + args.warmup + k, evs[k])
    t_end.record(stream)
    barrier()
    launches = wmix_b200.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_begin.elapsed_time(t_end)
    ns_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    post_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    mix_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    t_next = PRIME + args.warmup + args.steps

    # ---- end to end through the host-buffer C-ABI: pinned H2D of the tick, kernels, D2H of PCM + flags + bus
    h_out = [HostBuffer((S, FRAME), np.int16, device=local) for _ in range(2)]
    h_vad = [HostBuffer((S,), np.uint8, device=local) for _ in range(2)]
    h_bus = [HostBuffer((n_conf, FRAME), np.int32, device=local) for _ in range(2)]
    e2e_steps = max(10, min(args.steps, 100))

    def e2e_sync_step(t):
        # one C-ABI call: chunk-pipelined H2D, NS, AGC+VAD, bus, D2H of PCM + flags + bus; returns when the host has them
        eng.tick_host_bus(h_pool.array[t % R], h_out[0].array, h_vad[0].array, h_bus[0].array)

    for t in range(3):
        e2e_sync_step(t_next + t)
    barrier()
    t0 = time.perf_counter()
    for t in range(e2e_steps):
        e2e_sync_step(t_next + 3 + t)
    barrier()
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t_next += 3 + e2e_steps

    # the same call in its pipelined form (wmixb_tick_host_submit / _wait): ticks are fed back to back, two in flight, so
    # tick t+1's H2D overlaps tick t's last kernels and D2H.  Every step's copies are inside the timed region; the step's
    # result is on the host (and read) when its wait returns.  with_pcm = False: output selection, only the conference bus
    # and the speech flags come back (the tick's mix result) — the processed PCM stays on the device.
    sink = 0

    def e2e_pipelined(first, count, with_pcm=True):
        nonlocal sink
        for k in range(count):
            t = first + k
            eng.tick_host_submit(h_pool.array[t % R], h_out[k & 1].array if with_pcm else None, h_vad[k & 1].array, h_bus[k & 1].array)
            if k >= 1:
                eng.tick_host_wait()
                sink += int(h_bus[(k - 1) & 1].array[0, 0]) + int(h_vad[(k - 1) & 1].array[0])
        eng.tick_host_wait()
        sink += int(h_bus[(count - 1) & 1].array[0, 0])

    # A region of K ticks is short (K = 20: ~17 ms, one of which is the pipeline's drain) and a single host hiccup moves it by
    # several per cent, so the region is repeated and the MEDIAN region is reported (all regions are listed in the line).
    e2e, e2e_regions = {}, {}
    E2E_REGIONS = 5
    for name, with_pcm in (("full", True), ("bus_only", False)):
        e2e_pipelined(t_next, 6, with_pcm)
        t_next += 6
        regions = []
        for _ in range(E2E_REGIONS):
            barrier()
            t0 = time.perf_counter()
            e2e_pipelined(t_next, e2e_steps, with_pcm)
            barrier()
            regions.append((time.perf_counter() - t0) * 1e3 / e2e_steps)
            t_next += e2e_steps
        e2e_regions[name] = regions
        e2e[name] = sorted(regions)[E2E_REGIONS // 2]
    h2d_bytes = S * FRAME * 2
    d2h_full = S * FRAME * 2 + S + n_conf * FRAME * 4
    d2h_bus_only = S + n_conf * FRAME * 4
    # what the copy engines alone sustain for the same bytes, all ranks at once (no kernels): the ceiling of the host path
    barrier()
    h_sink = HostBuffer((d2h_full,), np.uint8, device=local)
    ceil_full = host_copy_ceiling(local, h_pool.array[0], h_sink.array, h2d_bytes, d2h_full, 20)
    barrier()
    ceil_bus = host_copy_ceiling(local, h_pool.array[0], h_sink.array, h2d_bytes, d2h_bus_only, 20)
    barrier()

    ms_step = ms_total / args.steps
    stats = torch.tensor([ms_step, e2e["full"], ns_ms, post_ms, mix_ms, e2e_sync_ms, e2e["bus_only"], ceil_full, ceil_bus],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_step, e2e_ms, ns_ms, post_ms, mix_ms, e2e_sync_ms, e2e_bus_ms, ceil_full, ceil_bus = [float(v) for v in stats.cpu()]
    state_mb = S * eng.state_bytes_per_stream() / 1e6
    eng.close()
    del d_pool, d_pcm

    conf5 = None
    if world > 1:
        conf5 = {"c1024": conf5_leg(torch, dist, world, rank, local, dev, 1024),
                 "c16": conf5_leg(torch, dist, world, rank, local, dev, 16)}
    full_load = None
    if not args.no_full_load:
        try:
            full_load = full_load_leg(torch, dev, local, args.full_load_streams)
            if world > 1:
                fl = torch.tensor([full_load["ms_per_tick"]], dtype=torch.float64, device=dev)
                dist.all_reduce(fl, op=dist.ReduceOp.MAX)
                full_load["ms_per_tick"] = float(fl.item())
                full_load["fits_10ms_tick"] = full_load["ms_per_tick"] <= 10.0
                full_load["streams_resident_whole_job"] = args.full_load_streams * world
        except Exception as ex:  # pragma: no cover
            full_load = {"failed": str(ex)}
    if rank == 0:
        peak, peak_src = peaks()
        total_streams = S * world
        achieved = S * NS_BYTES_PER_STREAM_TICK / (ns_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC,
            "value": total_streams * 10.0 / ms_step, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i16 (float NS with double transcendentals, integer AGC/VAD/mix)", "data": "synthetic",
            "config": {"workload": workload_text(S),
                       "streams_per_gpu": S, "tick_ms": 10, "prime_ticks": PRIME,
                       "parallelism": "streams sharded over the ranks, no collective on this path; the conference-bus exchange "
                                      "(config 5) is measured in `conf5` when n_gpus > 1",
                       "l2_policy": "per-tick working set (state %.0f MB + PCM) exceeds the 126 MB L2; inputs rotate over %d ticks"
                                    % (state_mb, R)},
            "realtime_headroom": 10.0 / ms_step,
            "kernel_ms": {"ns_kernel": ns_ms, "post_kernel(agc+vad)": post_ms, "bus_sum_kernel": mix_ms},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": S * NS_DRAM_TRAFFIC_PER_STREAM_NCU, "traffic_source": NS_TRAFFIC_SOURCE,
                         "kernel": "ns_cta_kernel<256, 8, 2>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": S * NS_BYTES_PER_STREAM_TICK,
                         "whole_tick_frac": S * CHAIN_BYTES_PER_STREAM_TICK / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": total_streams * 10.0 / e2e_ms, "unit": UNIT,
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_full, "steps": e2e_steps,
                    "regions": E2E_REGIONS, "ms_per_step_regions_rank0": [round(v, 4) for v in e2e_regions["full"]],
                    "regions_note": "ms_per_step = the median of `regions` timed regions of `steps` ticks each (max over ranks of the medians)",
                    "path": "wmixb_tick_host_submit / _wait, ticks fed back to back (two in flight): pinned host PCM in -> NS -> "
                            "AGC+VAD -> bus -> host PCM + VAD flags + bus, one chunk per CUDA stream (4); host buffers from "
                            "wmixb_host_alloc (pinned, placed on the GPU's NUMA node)",
                    "copy_ceiling_ms_per_step": ceil_full,
                    "frac_of_copy_ceiling": ceil_full / e2e_ms,
                    "copy_ceiling_note": "bare cudaMemcpyAsync of the same bytes up and down on two streams, all ranks at once, no kernels",
                    "bus_only": {"value": total_streams * 10.0 / e2e_bus_ms, "ms_per_step": e2e_bus_ms, "d2h_bytes_per_step": d2h_bus_only,
                                 "copy_ceiling_ms_per_step": ceil_bus, "frac_of_copy_ceiling": ceil_bus / e2e_bus_ms,
                                 "path": "same call with h_out = NULL (output selection): only the conference bus and the VAD flags "
                                         "return to the host"},
                    "sync_call_ms_per_step": e2e_sync_ms,
                    "sync_call_value": total_streams * 10.0 / e2e_sync_ms,
                    "sync_call_path": "wmixb_tick_host_bus: one blocking call per tick (pipeline drains at every tick boundary)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if conf5 is not None:
            line["conf5"] = conf5
        if full_load is not None:
            line["full_load"] = full_load
        if world == 1 and not args.no_config4:
            try:
                line["config4"] = config4_leg(torch, dev, local, peak)
            except Exception as ex:  # pragma: no cover
                line["config4"] = {"failed": str(ex)}
        if world == 1 and not args.no_nsx:
            try:
                line["nsx"] = nsx_leg(torch, dev, local, peak, S)
            except Exception as ex:  # pragma: no cover
                line["nsx"] = {"failed": str(ex)}
        if world == 1 and not args.no_offline:
            try:
                line["offline"] = offline_leg(torch, dev, local)
            except Exception as ex:  # pragma: no cover
                line["offline"] = {"failed": str(ex)}
        if not args.no_cpu_baseline and world == 1:
            try:
                leg = cpu_leg(2048, 40, prime=PRIME + min(max(args.warmup, 0), 300))
                line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["us_per_stream_tick_per_core"] = leg["us_per_stream_tick_per_core"]
            except Exception as ex:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": "", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=100_000, help="streams per GPU")
    ap.add_argument("--ring", type=int, default=8, help="distinct input ticks kept resident")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-full-load", action="store_true")
    ap.add_argument("--no-offline", action="store_true")
    ap.add_argument("--no-nsx", action="store_true")
    ap.add_argument("--full-load-streams", type=int, default=1_000_000)
    ap.add_argument("--ns-cfg", type=int, default=-1, help="experiment: NS kernel shape index (wmixb_set_tuning)")
    ap.add_argument("--post-occ", type=int, default=-1, help="experiment: AGC+VAD kernel shape (wmixb_set_tuning)")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE", help="experiment: any wmixb_set_tuning knob of the main engine")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
