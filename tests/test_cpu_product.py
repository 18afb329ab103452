"""CPU-side checks of the product: the C-ABI library builds, loads and exports every symbol
the headers declare (no compute without a GPU), the init-time tables match the reference's
literals, and the lane-loop emulation of the kernel bodies (tests/emu) agrees with the
reference bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests._emu import EmuChain, emu
from tests._oracle import P, RefChain, oracle, ref
from wmix_b200.synth import make_frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    import wmix_b200

    L = C.CDLL(wmix_b200.LIB_PATH)
    declared = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(\w+)\s*\([^;{]*\)\s*;", src))
    declared -= {"defined"}
    assert len(declared) > 40
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing


def test_headers_are_plain_c_and_link(tmp_path):
    """a C99 program (-pedantic) that includes every header under include/, links libwmix_b200.so and calls through the
    C-ABI: the boundary wmix's own C sources would use (tests/c/abi_from_c.c).  Without a GPU it must see WMIXB_ENODEV / NULL."""
    import subprocess

    import __graft_entry__ as g

    g.build()
    exe = str(tmp_path / "abi_from_c")
    libdir = os.path.join(ROOT, "wmix_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_from_c.c"), "-L", libdir, "-lwmix_b200", "-Wl,-rpath," + libdir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch

    run = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert "wmixb_create" in run.stdout                 # with a GPU the program reports success codes instead
    else:
        assert run.returncode == 0, run.stdout + run.stderr


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import wmix_b200

    with pytest.raises(wmix_b200.WmixError) as ei:
        wmix_b200.Engine(4, 16000)
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)
    assert not wmix_b200.lib().ns_init(1, 16000, None)


def test_ns_window_matches_reference_literals():
    hdr = "/tmp/wmix_ref_build/webrtc_cut/webrtc/modules/audio_processing/ns/windows_private.h"
    if not os.path.exists(hdr):
        pytest.skip("reference tarball not unpacked on this box")
    import wmix_b200

    src = open(hdr).read()
    for name, ana, block in (("kBlocks80w128", 128, 80), ("kBlocks160w256", 256, 160)):
        m = re.search(name + r"\[\d+\] = \{(.*?)\};", src, re.S)
        t = np.array([np.float32(v) for v in re.findall(r"\d+\.\d+", m.group(1))], dtype=np.float32)
        w = np.zeros(ana, np.float32)
        wmix_b200.lib().wmixb_ns_window(ana, block, w.ctypes.data)
        assert np.array_equal(t, w)


def test_agc_gain_table_product_vs_oracle_all_gains():
    import wmix_b200

    lib, L = wmix_b200.lib(), oracle()
    for comp in range(0, 91):
        for lim in (0, 1):
            at = lib.wmixb_agc_analog_target(comp)
            assert at == L.orc_agc_analog_target(comp)
            a = np.zeros(32, np.int32)
            b = np.zeros(32, np.int32)
            assert lib.wmixb_agc_gain_table(a.ctypes.data, comp, 0, lim, at) == L.orc_agc_gain_table(P(b), comp, 0, lim, at)
            assert np.array_equal(a, b), (comp, lim)


def test_emulated_g711_and_mix_full_domain():
    E, L = emu(), oracle()
    x = np.arange(-32768, 32768, dtype=np.int16)
    a = np.zeros(65536, np.uint8)
    u = np.zeros(65536, np.uint8)
    E.emu_g711(P(x), 65536, P(a), P(u))
    wa = np.zeros(65536, np.uint8)
    wu = np.zeros(65536, np.uint8)
    L.orc_PCM2G711a(P(x), P(wa), 131072)
    L.orc_PCM2G711u(P(x), P(wu), 131072)
    assert np.array_equal(a, wa) and np.array_equal(u, wu)
    c = np.arange(256, dtype=np.uint8)
    da = np.zeros(256, np.int16)
    du = np.zeros(256, np.int16)
    E.emu_g711_dec(P(c), 256, P(da), P(du))
    assert [int(v) for v in da] == [L.orc_alaw2linear(int(k)) for k in c]
    assert [int(v) for v in du] == [L.orc_ulaw2linear(int(k)) for k in c]
    rng = np.random.default_rng(0)
    for _ in range(2000):
        b, s, r = int(rng.integers(-32768, 32768)), int(rng.integers(-32768, 32768)), int(rng.integers(1, 17))
        ring = np.array([b], np.int16)
        L.orc_mix_same_format(P(ring), 1, 0, P(np.array([s], np.int16)), 1, r)
        assert E.emu_mix_step(b, s, r) == ring[0]


@pytest.mark.parametrize("freq", [16000, 8000])
@pytest.mark.parametrize("stage", ["vad", "agc", "ns", "chain"])
def test_emulated_kernel_bodies_vs_reference(freq, stage):
    """Same source as the sm_100a kernels, executed lane by lane on the CPU."""
    chk = ref() or oracle()
    prefix = "" if ref() is not None else "orc_"
    S, T = (5, 620) if stage in ("ns", "chain") else (6, 300)
    x = make_frames(S, freq, 0, T, seed=13)
    kw = dict(ns=stage in ("ns", "chain"), agc=stage in ("agc", "chain"), vad=stage in ("vad", "chain"))
    for s in range(S):
        pcm = np.ascontiguousarray(x[:, s, :]).reshape(-1)
        a = RefChain(chk, freq, prefix=prefix, **kw)
        b = EmuChain(freq, **kw)
        ya, yb = a.run(pcm), b.run(pcm)
        a.close()
        b.close()
        assert np.array_equal(ya, yb), (stage, freq, s)


@pytest.mark.parametrize("freq,workers,order", [(16000, 8, 0), (16000, 8, 1), (8000, 8, 0), (8000, 8, 1), (16000, 3, 1)])
def test_emulated_cta_ns_equals_single_warp_ns_and_reference(freq, workers, order):
    """The CTA-cooperative NS (ns_cta.cuh: worker warps + one reducer warp that walks every in-order sum, the scalar model
    and the Nyquist bin of all the CTA's streams) against the single-warp body (ns.cuh) it re-distributes: outputs, the
    whole per-stream record and the feature histograms, bit for bit, over 720 frames (past the start-up model at 50, the
    gain map at 200 and the first histogram re-learning at 500), with an all-zero cohort stream, a stream that joins late
    and an idle worker slot.  One stream is also checked against the reference itself.  `order`: the reducer's deferred parts
    (r_seg1b, r_seg2b) run concurrently with the workers' next segment on the device — here either side goes first."""
    import ctypes as C

    E = emu()
    n = freq // 100
    T = 720
    x = make_frames(workers, freq, 0, T, seed=29)
    if workers > 1:
        x[:, 1, :] = 0                                   # a zero stream: energy == 0 early-outs every frame
    if workers > 2:
        x[100:140, 2, :] = 0                             # zero frames in the middle of a live stream
    h = C.c_void_p(E.emu_nscta_create(freq, workers))
    E.emu_nscta_order(h, order)
    singles = [C.c_void_p(E.emu_ns_create(freq)) for _ in range(workers)]
    rec_floats = E.emu_ns_rec_floats(freq)
    live = np.ones(workers, np.uint8)
    out = np.zeros((workers, n), np.int16)
    want = np.zeros((workers, n), np.int16)
    all_out = np.zeros((T, workers, n), np.int16)
    for t in range(T):
        live[:] = 1
        if workers > 3 and t < 33:
            live[3] = 0                                  # worker 3's stream starts 33 frames late (different frame index)
        frame_in = np.ascontiguousarray(x[t])
        out[:] = 0
        E.emu_nscta_frame(h, P(frame_in), P(out), P(live))
        for j in range(workers):
            if not live[j]:
                continue
            buf = frame_in[j].copy()
            E.emu_ns_frame(singles[j], P(buf), P(buf))
            want[j] = buf
        ok = live.astype(bool)
        assert np.array_equal(out[ok], want[ok]), (freq, t, np.argwhere(out[ok] != want[ok])[:4])
        all_out[t] = out
        if t in (0, 1, 49, 50, 199, 200, 201, 499, 500, 501, T - 1):
            for j in range(workers):
                a = np.ctypeslib.as_array(E.emu_nscta_record(h, j), (rec_floats,)).view(np.uint32)
                b = np.ctypeslib.as_array(E.emu_ns_record(singles[j]), (rec_floats,)).view(np.uint32)
                assert np.array_equal(a, b), (freq, t, j, np.argwhere(a != b)[:8].ravel())
                ha = np.ctypeslib.as_array(E.emu_nscta_hist(h, j), (3000,))
                hb = np.ctypeslib.as_array(E.emu_ns_hist(singles[j]), (3000,))
                assert np.array_equal(ha, hb), (freq, t, j)
    # and one stream against the reference (or its restatement)
    chk = ref() or oracle()
    a = RefChain(chk, freq, prefix="" if ref() is not None else "orc_", ns=True, agc=False, vad=False)
    ya = a.run(np.ascontiguousarray(x[:, 0, :]).reshape(-1))
    a.close()
    assert np.array_equal(ya, all_out[:, 0, :].reshape(-1))
    E.emu_nscta_destroy(h)
    for s_ in singles:
        E.emu_ns_destroy(s_)


@pytest.mark.parametrize("freq", [16000, 8000])
def test_emulated_vad_20ms_packets_vs_reference(freq):
    """wmix itself creates its VAD with intervalMs = 20 (R:src/wmix.c:703): 20 ms packets and the 20 ms threshold
    column (T:.../vad/vad_core.c:149-164).  Kernel body (lane-loop emulation) against the reference handle."""
    import ctypes as C

    chk = ref() or oracle()
    n = freq // 50
    x = make_frames(6, freq, 0, 400, seed=17)                      # 400 ticks = 200 packets per stream
    E = emu()
    E.emu_int_create.restype = C.c_void_p
    for s in range(6):
        pcm = np.ascontiguousarray(x[:, s, :]).reshape(-1, n)
        if ref() is not None:
            chk.vad_init.restype = C.c_void_p
            h = C.c_void_p(chk.vad_init(1, freq, 20, None))
            proc, rel = chk.vad_process, chk.vad_release
        else:
            h = C.c_void_p(chk.orc_vad_init(1, freq, 20))
            proc, rel = chk.orc_vad_process, chk.orc_vad_release
        e = C.c_void_p(E.emu_int_create(freq, 5, 3))
        for k in range(len(pcm)):
            a, b = pcm[k].copy(), pcm[k].copy()
            proc(h, P(a), n)
            E.emu_vad_frame20(e, P(b), 3)
            assert np.array_equal(a, b), (freq, s, k)
        rel(h)
        E.emu_int_destroy(e)


@pytest.mark.parametrize("freq", [16000, 8000])
def test_emulated_ns_stereo_high_band_vs_oracle(freq):
    """ns_init(2, ..): the right channel rides through WebRtcNs as a second band (R:src/webrtc.c:624-636;
    T:.../ns/ns_core.c:1214-1261, :1361-1414).  Kernel body frame<ANA, true> (lane-loop emulation) against the
    oracle, which tests/test_oracle_pin.py pins to the reference for this case; pair (1, 2) has an all-zero left
    channel (the energy == 0 early-out also flushes the high band)."""
    import ctypes as C

    E, L = emu(), oracle()
    n = freq // 100
    T = 300
    x = make_frames(4, freq, 0, T, seed=31)
    for a, b in ((0, 3), (1, 2), (2, 2)):
        h = C.c_void_p(L.orc_ns_init(2, freq))
        e = C.c_void_p(E.emu_ns_create(freq))
        for t in range(T):
            st = np.empty(2 * n, np.int16)
            st[0::2], st[1::2] = x[t, a], x[t, b]
            want = np.zeros(2 * n, np.int16)
            L.orc_ns_process(h, P(st), P(want), n)
            lo, hi = x[t, a].copy(), x[t, b].copy()
            E.emu_ns_frame_hb(e, P(lo), P(lo), P(hi), P(hi))          # in place, as wmix calls it
            assert np.array_equal(lo, want[0::2]) and np.array_equal(hi, want[1::2]), (freq, a, b, t)
        L.orc_ns_release(h)
        E.emu_ns_destroy(e)


def test_vad_32khz_packets_match_oracle():
    """vad_init(chn, 32000, ..): 320-sample packets through CalcVad32khz (T:.../vad/vad_core.c:623-643) — the kernel body of
    wmixb_vad32_device against the oracle (pinned to the reference at 32 kHz in tests/test_oracle_pin.py)"""
    import ctypes as C

    E, chk = emu(), oracle()
    x = make_frames(3, 16000, 0, 240, seed=29)                       # 160-sample rows; pairs of rows = one 320-sample packet
    for s in range(3):
        pcm = np.ascontiguousarray(x[:, s]).reshape(-1, 320)
        h = C.c_void_p(chk.orc_vad_init(1, 32000, 10))
        e = C.c_void_p(E.emu_int_create(16000, 5, 3))
        for k in range(len(pcm)):
            a, b = pcm[k].copy(), pcm[k].copy()
            chk.orc_vad_process(h, P(a), 320)
            E.emu_vad_frame32(e, P(b))
            assert np.array_equal(a, b), (s, k)
        chk.orc_vad_release(h)
        E.emu_int_destroy(e)


def test_ns_counter_division_is_exact():
    """ns::div_by_counter (reciprocal + exact residual + one correction) replaces `x / (counter + 1)` in the NS quantile
    trackers (T:.../ns/ns_core.c:233-249); it must BE the IEEE quotient: every divisor 1..201 on a significand grid, and
    every one of the 2^23 significands for a handful of divisors."""
    L = emu()
    assert L.emu_div_by_counter_mismatches(1, 201, 16) == 0
    for d in (3, 67, 134, 199, 200, 201):
        assert L.emu_div_by_counter_mismatches(d, d, 1) == 0


def test_synth_streams_are_reproducible_and_tick_addressable():
    a = make_frames(70, 16000, 0, 30, seed=3)
    b = make_frames(70, 16000, 10, 5, seed=3)
    assert np.array_equal(a[10:15], b)
    assert not a[:, 1].any() and np.abs(a[:, 2]).min() == 32767


# ---------------------------------------------------------------- AEC (lane-loop emulation of aec.cuh)
AEC_MAX_ABS = 1          # powf / cosf / sinf are evaluated in double and rounded: a last-bit difference of the
AEC_MIN_EQUAL = 0.9999   # suppression gain or comfort noise can move an output sample by one LSB, never the state


def _aec_emu_vs_oracle(freq, n, ims, T, S, depth, delay=0, bursts=False, seed=3):
    from tests._oracle import AecRef
    from wmix_b200.synth import make_aec_pairs

    E, L = emu(), oracle()
    E.emu_aec_create.restype = C.c_void_p
    far, near = make_aec_pairs(S, freq, 0, T, seed=seed)
    far = far.transpose(1, 0, 2).reshape(S, -1, n)
    near = near.transpose(1, 0, 2).reshape(S, -1, n)
    bad = tot = worst = flags = 0
    for s in range(S):
        a = AecRef(L, freq, ims, "orc_")
        b = C.c_void_p(E.emu_aec_create(freq, depth))
        t, nt = 0, far.shape[1]
        while t < nt:
            grp = min(nt - t, 1 + (t * 7 + s) % 3) if bursts else 1
            outs = []
            if bursts:
                for g in range(grp):
                    a.set_far(far[s, t + g])
                    E.emu_aec_tick(b, P(np.ascontiguousarray(far[s, t + g])), None, None, n, 0)
                for g in range(grp):
                    ya, _ = a.process(near[s, t + g], delay)
                    yb = np.zeros(n, np.int16)
                    E.emu_aec_tick(b, None, P(np.ascontiguousarray(near[s, t + g])), P(yb), n, delay)
                    outs.append((ya, yb))
            else:
                ya, _ = a.process2(far[s, t], near[s, t], delay)
                yb = np.zeros(n, np.int16)
                E.emu_aec_tick(b, P(np.ascontiguousarray(far[s, t])), P(np.ascontiguousarray(near[s, t])), P(yb), n, delay)
                outs.append((ya, yb))
            for ya, yb in outs:
                d = np.abs(ya.astype(np.int32) - yb)
                bad += int((d > 0).sum())
                tot += n
                worst = max(worst, int(d.max()))
            t += grp
        flags |= E.emu_aec_error(b)
        a.close()
        E.emu_aec_destroy(b)
    return bad, tot, worst, flags


@pytest.mark.parametrize("freq,n,ims,T,S,depth,delay,bursts", [
    (8000, 80, 10, 700, 8, 32, 0, False),       # config 4's rate and cadence; includes the zero-far / square cohorts
    (16000, 160, 10, 400, 4, 32, 0, False),
    (8000, 160, 20, 300, 3, 32, 40, False),     # wmix's own 20 ms packets
    (8000, 80, 10, 500, 3, 32, 120, True),      # far / near in bursts, non-zero reported delay
    (16000, 160, 10, 300, 2, 252, 0, True),     # the handle API's full-depth ring
    (8000, 80, 10, 300, 2, 252, 480, False),    # a 480 ms reported delay: the far read pointer rewinds ~60 partitions
])
def test_emulated_aec_vs_oracle(freq, n, ims, T, S, depth, delay, bursts):
    bad, tot, worst, flags = _aec_emu_vs_oracle(freq, n, ims, T, S, depth, delay, bursts)
    print("[aec emu %d Hz n=%d] mismatching %d / %d, max |diff| %d, flags %d" % (freq, n, bad, tot, worst, flags))
    assert flags == 0
    assert worst <= AEC_MAX_ABS and bad <= (1.0 - AEC_MIN_EQUAL) * tot


def test_emulated_aec_rdft_and_tables():
    import wmix_b200

    E, L = emu(), oracle()
    rng = np.random.default_rng(2)
    for k in range(100):
        a = (rng.standard_normal(128) * 10 ** rng.uniform(-3, 4)).astype(np.float32)
        for inv in (0, 1):
            x, y = a.copy(), a.copy()
            L.orc_aec_rdft(P(x), inv)
            E.emu_aec_rdft(P(y), inv)
            assert np.array_equal(x.view(np.int32), y.view(np.int32))
    assert wmix_b200  # the product's tables are pinned through the GPU parity tests


def test_aec_far_depth_flag_is_loud():
    """A far-end backlog deeper than the configured history must raise the sticky flag, not corrupt silently."""
    E = emu()
    E.emu_aec_create.restype = C.c_void_p
    h = C.c_void_p(E.emu_aec_create(8000, 8))
    z = np.zeros(80, np.int16)
    o = np.zeros(80, np.int16)
    for t in range(40):                       # start-up: process normally
        E.emu_aec_tick(h, P(z), P(z), P(o), 80, 0)
    assert E.emu_aec_error(h) == 0
    for t in range(30):                       # 30 far frames without a near frame: 37 partitions of backlog
        E.emu_aec_tick(h, P(z), None, None, 80, 0)
    E.emu_aec_tick(h, None, P(z), P(o), 80, 0)
    assert E.emu_aec_error(h) & 1
    E.emu_aec_destroy(h)


def test_zoom_host_routing_table_vs_oracle():
    """wmix_len_of_in / wmix_len_of_out (host arithmetic of the drop-in) and the gather table that the CUDA kernel
    applies, against the oracle's sample-moving walk: routing known samples through the table must give the oracle's
    output.  (Needs no GPU: the table is init-time host logic, exported for exactly this check.)"""
    import ctypes as C

    from tests.test_oracle_pin import ZOOM_CASES

    E, L = emu(), oracle()
    for f in (E.emu_zoom_map, E.emu_len_of_out, E.emu_len_of_in, L.orc_pcm_zoom, L.orc_len_of_out, L.orc_len_of_in):
        f.restype = C.c_uint32
    rng = np.random.default_rng(3)
    for ic, ifr, oc, ofr in ZOOM_CASES:
        for in_bytes in (2 * ic, 320 * ic, 640, 3528, 2 * ic * 777):
            x = rng.integers(-32768, 32768, in_bytes // 2).astype(np.int16)
            cap = 16 * in_bytes * max(1, ofr // ifr + 1) + 64
            want = np.zeros(cap, np.int16)
            nb = L.orc_pcm_zoom(ic, ifr, P(x.copy()), in_bytes, oc, ofr, P(want))
            m = np.zeros(cap, np.int32)
            n = E.emu_zoom_map(ic, ifr, in_bytes, oc, ofr, P(m))
            assert 2 * n == nb, (ic, ifr, oc, ofr, in_bytes)
            assert np.array_equal(x[m[:n]], want[:n])
            assert E.emu_len_of_out(ic, ifr, in_bytes, oc, ofr) == L.orc_len_of_out(ic, ifr, in_bytes, oc, ofr)
            assert E.emu_len_of_in(ic, ifr, oc, ofr, in_bytes) == L.orc_len_of_in(ic, ifr, oc, ofr, in_bytes)


MIX_RESAMPLE_CASES = [(8000, 1), (8000, 2), (11025, 1), (12000, 2), (16000, 2), (22050, 1), (32000, 2), (44100, 2),
                      (48000, 1), (15999, 1), (16001, 2), (300, 1), (65535, 1)]


def apply_mix_plan(ring, pos, src, chn, m, ramp, d):
    """numpy model of mix_plan_kernel for ONE source (float32 arithmetic step by step, like the kernel)"""
    f32 = np.float32
    for i in range(len(m)):
        v = int(src[m[i]])
        if ramp[i]:
            n, k = int(ramp[i]) >> 8, int(ramp[i]) & 0xFF
            step = f32(int(src[m[i] + chn]) - v) / f32(n)
            run = step
            for _ in range(1, k):
                run = f32(run + step)
            v = int(np.trunc(f32(f32(v) + run)))
        q = -(-v // d) if v < 0 else v // d      # C division truncates toward zero
        p = (pos + i) % len(ring)
        ring[p] = max(-32768, min(32767, int(ring[p]) + q))


def test_mix_resample_host_plan_vs_oracle():
    """the host-built plan of wmixb_mix_load_plan_device (which source sample / which ramp step feeds every bus
    sample) against the oracle's sequential walk of wmix_load_data's resampling branches (R:src/wmix.c:1704-1939)"""
    import ctypes as C

    E, L = emu(), oracle()
    E.emu_mix_plan.restype = C.c_uint32
    L.orc_mix_resample.restype = C.c_uint32
    L.orc_mix_resample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8,
                                   C.c_uint16, C.c_uint8, C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(21)
    for mix_freq in (16000, 8000):
        for k, (freq, chn) in enumerate(MIX_RESAMPLE_CASES):
            if freq == mix_freq and chn == 1:
                continue
            frames = 257 if freq >= 1000 else 30
            src = rng.integers(-32768, 32768, frames * chn).astype(np.int16)
            if k % 3 == 0:
                src[:] = np.where(rng.random(src.size) < 0.5, 32767, -32768)
            nbytes = src.nbytes
            cap = frames * (mix_freq // freq + 2) + 8
            m, ramp = np.zeros(cap, np.int32), np.zeros(cap, np.uint16)
            n = E.emu_mix_plan(chn, freq, nbytes, mix_freq, P(m), P(ramp))
            if n == 0xFFFFFFFF:
                assert mix_freq // freq >= 63
                continue
            d = (1, 3, 16)[k % 3]
            ring_len = n + 37
            ring0 = rng.integers(-32768, 32768, ring_len).astype(np.int16)
            want, got = ring0.copy(), ring0.copy()
            wr = C.c_uint32(0)
            pos = L.orc_mix_resample(P(want), ring_len, ring_len - 20, P(src), nbytes, freq, chn, mix_freq, d, C.byref(wr))
            assert wr.value == n and pos == (ring_len - 20 + n) % ring_len, (freq, chn)
            apply_mix_plan(got, ring_len - 20, src, chn, m[:n], ramp[:n], d)
            assert np.array_equal(want, got), (mix_freq, freq, chn)
    # a ratio the reference's 64-entry ramp buffer cannot hold is refused
    assert E.emu_mix_plan(1, 200, 200, 16000, None, None) == 0xFFFFFFFF


def test_play_fifo_slot_arithmetic_vs_oracle():
    """playPkgBuff_get's slot choice (R:src/wmix.c:496-509) as the record tick computes it on the host, against the oracle
    (itself pinned to the reference's playPkgBuff_add / _get) for every ring size, write index and delay"""
    E, L = emu(), oracle()
    for n_pkg in range(2, 40):
        for count in range(n_pkg):
            for d in range(0, n_pkg + 6):
                assert E.emu_play_fifo_slot(count, n_pkg, d) == L.orc_play_fifo_slot(count, n_pkg, d), (n_pkg, count, d)


def test_rtp_host_header_helpers_vs_oracle():
    """wmixb_rtp_write_header / wmixb_rtp_read_header (host byte shuffling of the C-ABI) against the oracle"""
    import ctypes as C

    import wmix_b200

    lib, L = wmix_b200.lib(), oracle()
    rng = np.random.default_rng(9)
    for _ in range(500):
        cc, x, p, v = int(rng.integers(0, 16)), int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 4))
        pt, m = int(rng.integers(0, 128)), int(rng.integers(0, 2))
        seq, ts, ssrc = int(rng.integers(0, 65536)), int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32))
        a, b = np.zeros(12, np.uint8), np.zeros(12, np.uint8)
        L.orc_rtp_header_bytes(P(a), cc, x, p, v, pt, m, seq, ts, ssrc)
        lib.wmixb_rtp_write_header(b.ctypes.data, cc | (x << 4) | (p << 5) | (v << 6), m, pt, seq, ts, ssrc)
        assert np.array_equal(a, b)
        meta = np.zeros(16, np.uint8)
        lib.wmixb_rtp_read_header(a.ctypes.data, meta.ctypes.data)
        f = meta.view(np.uint32)
        assert f[0] == ts and f[1] == ssrc and meta[8:10].view(np.uint16)[0] == seq
        assert meta[10] == pt and meta[11] == m and meta[12] == a[0] and meta[13] == int(v == 2 and pt in (0, 8))


def test_vad_minimum_tracker_on_packed_lists_vs_oracle():
    """vad::find_minimum works on the record's packed (age, value) pairs with funnel shifts; against the oracle's list code
    (pinned to the reference's WebRtcVad_FindMinimum KAT) over feature sequences built to hit its corners: all sixteen initial
    entries expiring in one frame, the entry behind an expired one skipping its ageing, expiry at even and odd places,
    insertions at every place, values at and above the 10000 sentinel, long stretches without any insertion"""
    import ctypes as C

    E, L = emu(), oracle()
    rng = np.random.default_rng(12)
    from tests.test_oracle_pin import VadCore

    E.emu_vad_minimum_run.restype = None
    L.orc_vad_find_minimum.restype = C.c_int16

    def sequences():
        yield np.full(230, 12000, np.int16)                                   # nothing ever inserted: the 16 initial entries expire together
        yield np.concatenate([np.arange(3000, 3016), np.full(240, 9000)]).astype(np.int16)       # 16 consecutive inserts, then they expire one per frame
        yield np.concatenate([np.arange(3016, 3000, -1), np.full(240, 9999)]).astype(np.int16)   # always at place 0
        yield np.concatenate([np.full(40, 10000), np.arange(100, 140), np.full(150, 10001)]).astype(np.int16)
        for _ in range(40):
            n = int(rng.integers(150, 700))
            base = rng.integers(-2000, 11000, size=n)
            hold = rng.integers(0, 2, size=n).astype(bool)                     # stretches of large values let entries reach 100
            seq = np.where(hold, rng.integers(9000, 12000, size=n), base)
            yield seq.astype(np.int16)

    for k, feats in enumerate(sequences()):
        ch = k % 6
        core = VadCore()
        L.orc_vad_core_init(C.byref(core), 3)
        want = np.zeros(len(feats), np.int16)
        for t, f in enumerate(feats):
            core.frame_counter = t
            want[t] = L.orc_vad_find_minimum(C.byref(core), int(f), ch)
        got = np.zeros(len(feats), np.int16)
        words = np.zeros(17, np.int32)
        E.emu_vad_minimum_run(P(np.ascontiguousarray(feats)), len(feats), ch, P(got), P(words))
        assert np.array_equal(got, want), (k, int(np.nonzero(got != want)[0][0]))
        ages = np.array(core.age[ch * 16:(ch + 1) * 16], np.int16)
        lows = np.array(core.low_value[ch * 16:(ch + 1) * 16], np.int16)
        assert np.array_equal(words[:8].view(np.int16), ages), k
        assert np.array_equal(words[8:16].view(np.int16), lows), k
