"""GPU parity of the fixed-point suppressor (ns_core = 1; nsx.cuh / nsx_kernel) through the C-ABI.  Integer work: the bar
is BIT-EXACT, outputs and complete per-stream state, against oracle/orc_nsx.c, the compiled reference's WebRtcNsx_* when it
travelled with the snapshot, and the committed fixtures of tests/golden/nsx.json (recorded from the unmodified reference)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import wmix_b200  # noqa: E402
from tests._oracle import NsxCore, P, fnv1a64, nsx_quiet_streams, oracle, ref  # noqa: E402
from tests.test_nsx_oracle_pin import check, oracle_core_run  # noqa: E402
from wmix_b200 import AGC, NS, VAD  # noqa: E402
from wmix_b200.synth import make_frames  # noqa: E402

DEV = "cuda:0"
G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "nsx.json")))


def run_gpu_nsx(x, freq, policy=2, stages=NS, offline=0, tuning=None):
    """x int16 [T, S, L] -> out [T, S, L] through a batched engine with the fixed-point core"""
    T, S, L = x.shape
    eng = wmix_b200.Engine(S, freq, stages=stages, ns_policy=policy, ns_core=1)
    for k, v in (tuning or {}).items():
        eng.set_tuning(k, v)
    out = np.empty_like(x)
    if offline:
        assert T % offline == 0
        for t0 in range(0, T, offline):
            d_in = torch.from_numpy(np.ascontiguousarray(x[t0:t0 + offline].transpose(1, 0, 2))).to(DEV)
            d_out = torch.empty_like(d_in)
            eng.offline_device(d_in, d_out, offline)
            out[t0:t0 + offline] = d_out.cpu().numpy().transpose(1, 0, 2)
    else:
        d = torch.empty((S, L), dtype=torch.int16, device=DEV)
        for t in range(T):
            d.copy_(torch.from_numpy(x[t]))
            eng.tick_device(d, d)                                   # in place, like wmix
            out[t] = d.cpu().numpy()
    eng.close()
    return out


@pytest.mark.parametrize("freq", [16000, 8000])
@pytest.mark.parametrize("policy", [2, 0, 1, 3])
def test_nsx_golden_fixtures_through_the_gpu(freq, policy):
    """the reference's own outputs (1100 ticks: start-up model, gain map from frame 200, two threshold re-learnings)"""
    d = G["core"]["synth_%d_p%d" % (freq, policy)]
    x = make_frames(d["n_streams"], freq, 0, d["n_ticks"], seed=d["seed"])
    check(d, run_gpu_nsx(x, freq, policy))


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_vs_checkers_70_streams(freq):
    """a ragged batch (70 streams: partial CTA, every cohort of the synthetic set) over 650 ticks"""
    x = make_frames(70, freq, 0, 650, seed=41)
    got = run_gpu_nsx(x, freq)
    want = oracle_core_run(freq, x[:, :24])
    assert np.array_equal(got[:, :24], want)
    R = ref()
    if R is not None:
        for s in (0, 1, 2, 3, 17, 64, 65, 69):
            c = NsxCore(R, freq, 2)
            for t in range(650):
                assert np.array_equal(c.frame(x[t, s]), got[t, s]), (s, t)
            c.close()


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_quiet_gapped_saturated(freq):
    """near-silent streams (the modulo-32 shifts of the x86 reference), 30 all-zero frames (the zero-input path), full scale"""
    check(G["core"]["quiet_%d" % freq], run_gpu_nsx(nsx_quiet_streams(freq), freq))


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_state_word_for_word(freq):
    """complete per-stream state after 40, 230 and 600 ticks against the oracle's canonical dump (the emulator's
    record-to-canonical mapping is reused through wmixb_get_stream_state)"""
    from tests._emu import emu
    E = emu()
    O = oracle()
    O.orc_nsx_init_policy.restype = C.c_void_p
    S, L = 40, freq // 100
    x = make_frames(S, freq, 0, 600, seed=19)
    eng = wmix_b200.Engine(S, freq, stages=NS, ns_core=1)
    hs = [C.c_void_p(O.orc_nsx_init_policy(1, freq, 2)) for _ in range(S)]
    d = torch.empty((S, L), dtype=torch.int16, device=DEV)
    words = E.emu_nsx_rec_words(freq)
    for t in range(600):
        d.copy_(torch.from_numpy(x[t]))
        eng.tick_device(d, d)
        for s in range(S):
            y = np.zeros(L, np.int16)
            O.orc_nsx_process(hs[s], P(x[t, s].copy()), P(y), L)
        if t + 1 in (40, 230, 600):
            for s in range(S):
                raw = eng.get_state(s)
                rec = np.ascontiguousarray(raw[:words * 4]).view(np.uint32)
                hist = np.ascontiguousarray(raw[words * 4:words * 4 + 6000]).view(np.int16)
                a = np.zeros(4096, np.int32)
                b = np.zeros(4096, np.int32)
                n = O.orc_nsx_state(hs[s], P(a), 4096)
                assert E.emu_nsx_canonical(freq, P(rec), P(b)) == n
                assert np.array_equal(a[:n], b[:n]), (t, s, np.nonzero(a[:n] != b[:n])[0][:8])
                assert 0 <= int(hist.sum()) <= 3 * 511                          # at most three increments per counted frame of a window
    for h in hs:
        O.orc_nsx_release(h)
    eng.close()


def test_nsx_offline_mode_and_every_launch_shape_agree():
    """K frames per launch == K ticks; every compiled shape of the kernel (registers capped at 128 / 80 / 64, 4 .. 32
    warps per CTA) gives the same bits"""
    x = make_frames(70, 16000, 0, 240, seed=57)
    base = run_gpu_nsx(x, 16000)
    assert np.array_equal(base, run_gpu_nsx(x, 16000, offline=60))
    for cfg in range(1, 10):
        assert np.array_equal(base, run_gpu_nsx(x, 16000, tuning={"nsx_cfg": cfg})), cfg
    # with and without the frame-start barrier, ragged last CTA (70 streams), all-zero streams in the batch
    for cfg in (4, 7):
        assert np.array_equal(base, run_gpu_nsx(x, 16000, tuning={"nsx_sync": 0, "nsx_cfg": cfg})), cfg


def test_nsx_offline_across_a_threshold_relearning():
    """60 frames per launch over 660 frames: the 512-frame re-learning reads histograms whose last increments were issued as
    fire-and-forget REDs earlier in the same launch"""
    x = make_frames(40, 16000, 0, 660, seed=59)
    assert np.array_equal(oracle_core_run(16000, x[:, :10]), run_gpu_nsx(x, 16000, offline=60)[:, :10])
    x8 = make_frames(40, 8000, 0, 660, seed=60)
    assert np.array_equal(oracle_core_run(8000, x8[:, :10]), run_gpu_nsx(x8, 8000, offline=110)[:, :10])


def test_nsx_chain_with_agc_and_vad_and_config1_wav():
    """NSX -> AGC -> VAD (the record chain with the switch thrown) against the oracle chain; config 1's wav through NSX"""
    from tests._oracle import RefChain
    O = oracle()
    O.orc_nsx_init_policy.restype = C.c_void_p
    x = make_frames(12, 16000, 0, 400, seed=61)
    got = run_gpu_nsx(x, 16000, stages=NS | AGC | VAD)
    nsx = oracle_core_run(16000, x)
    for s in range(12):
        c = RefChain(O, 16000, ns=False, prefix="orc_")
        for t in range(400):
            assert np.array_equal(c.frame(nsx[t, s]), got[t, s]), (s, t)
        c.close()
    wav = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "config1_in_20s.s16"), np.int16).reshape(-1, 1, 80)
    check(G["config1_nsx"], run_gpu_nsx(wav, 8000).reshape(-1, 80))


def test_nsx_full_size_replication_property():
    """100 000 streams (BASELINE config 3's size): replicated inputs give replicated outputs equal to the small run"""
    S, base, T = 100000, 50, 12
    x = make_frames(base, 16000, 0, T, seed=67)
    small = run_gpu_nsx(x, 16000)
    eng = wmix_b200.Engine(S, 16000, stages=NS, ns_core=1)
    d = torch.empty((S, 160), dtype=torch.int16, device=DEV)
    for t in range(T):
        d.copy_(torch.from_numpy(x[t]).to(DEV).repeat(S // base, 1))
        eng.tick_device(d, d)
        y = d.view(S // base, base, 160)
        assert bool((y == y[0:1]).all())
        assert np.array_equal(y[0].cpu().numpy(), small[t])
    eng.close()


@pytest.mark.parametrize("freq", [8000, 16000, 32000])
@pytest.mark.parametrize("chn", [1, 2])
def test_nsx_dropin_handles_with_the_switch_thrown(chn, freq):
    """ns_init / ns_process after wmixb_set_default_ns_core(1) — the reference's `#define MAKE_WEBRTC_NSX` — against the
    outputs of the reference built with that define (tests/golden/nsx.json), the oracle, and the reference itself when it
    is here: mono and stereo (right channel as the second band), 8 / 16 / 32 kHz (320-sample packets)."""
    from tests._oracle import nsx_handle_run, ref_nsx
    lib = wmix_b200.lib()
    d = G["handle"]["%d_%d" % (chn, freq)]
    x = make_frames(chn, freq, 0, d["n_ticks"], seed=d["seed"])
    pcm = np.ascontiguousarray(x.transpose(0, 2, 1).reshape(d["n_ticks"], -1))
    assert lib.wmixb_set_default_ns_core(1) == 0 and lib.wmixb_default_ns_core() == 1
    try:
        got = nsx_handle_run(lib, "", chn, freq, pcm)
    finally:
        assert lib.wmixb_set_default_ns_core(0) == 0
    check(d, got)
    assert np.array_equal(got, nsx_handle_run(oracle(), "orc_nsx", chn, freq, pcm))
    if ref_nsx() is not None:
        assert np.array_equal(got, nsx_handle_run(ref_nsx(), "", chn, freq, pcm))
    assert lib.wmixb_set_default_ns_core(2) != 0                      # only 0 and 1 exist


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_second_band_batched(freq):
    """wmixb_ns2_device on an engine with ns_core = 1 and ns_high_band = 1 against the reference's two-band fixtures"""
    d = G["core"]["stereo_%d" % freq]
    lo = make_frames(4, freq, 0, 1100, seed=d["seed"])
    hi = make_frames(4, freq, 0, 1100, seed=d["seed_hb"])
    L = freq // 100
    eng = wmix_b200.Engine(4, freq, stages=NS, ns_core=1, ns_high_band=1)
    a = torch.empty((4, L), dtype=torch.int16, device=DEV)
    b = torch.empty((4, L), dtype=torch.int16, device=DEV)
    out_lo, out_hi = np.empty_like(lo), np.empty_like(hi)
    st = torch.cuda.current_stream().cuda_stream
    for t in range(1100):
        a.copy_(torch.from_numpy(lo[t]))
        b.copy_(torch.from_numpy(hi[t]))
        assert eng.L.wmixb_ns2_device(eng.h, a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), st) == 0     # in place
        out_lo[t], out_hi[t] = a.cpu().numpy(), b.cpu().numpy()
    eng.close()
    check(d["lo"], out_lo)
    check(d["hi"], out_hi)


@pytest.mark.parametrize("ns_core", [0, 1])
def test_32khz_batched_engine_like_the_reference_handles(ns_core):
    """wmixb_create(freq = 32000): 320-sample rows through the 16 kHz cores exactly as ns_init / agc_init / vad_init(.., 32000, ..)
    of the reference treat them — NS on the first 160 samples with zeros behind, AGC on two 5 ms packets, VAD on the
    320-sample packet — for both suppressors, device and host ticks, stage by stage and chained, against the oracle's 32 kHz
    handles (pinned to the reference's at that rate, tests/test_oracle_pin.py)."""
    O = oracle()
    O.orc_nsx_init.restype = C.c_void_p
    S, T = 40, 230
    x = make_frames(S, 16000, 0, 2 * T, seed=87)                         # [2T, S, 160] -> packets of 320
    pk = np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(S, T, 320)
    for stages in (NS, AGC, VAD, NS | AGC | VAD):
        eng = wmix_b200.Engine(S, 32000, stages=stages, ns_core=ns_core)
        assert eng.L.wmixb_frame_len(eng.h) == 320
        hs = []
        for s in range(S):
            ns = C.c_void_p((O.orc_nsx_init if ns_core else O.orc_ns_init)(1, 32000)) if stages & NS else None
            agc = C.c_void_p(O.orc_agc_init(1, 32000, 10, 5)) if stages & AGC else None
            vad = C.c_void_p(O.orc_vad_init(1, 32000, 10)) if stages & VAD else None
            hs.append((ns, agc, vad))
        d = torch.empty((S, 320), dtype=torch.int16, device=DEV)
        d_v = torch.zeros((S,), dtype=torch.uint8, device=DEV)
        h_out = np.empty((S, 320), np.int16)
        for t in range(T):
            want = pk[:, t].copy()
            for s, (ns, agc, vad) in enumerate(hs):
                f = want[s]
                if ns:
                    (O.orc_nsx_process if ns_core else O.orc_ns_process)(ns, P(f), P(f), 320)
                if agc:
                    assert O.orc_agc_process(agc, P(f), P(f), 320) == 0
                if vad:
                    O.orc_vad_process(vad, P(f), 320)
            if t % 2 == 0:
                d.copy_(torch.from_numpy(np.ascontiguousarray(pk[:, t])))
                eng.tick_device(d, d, d_v)
                got = d.cpu().numpy()
            else:
                eng.tick_host(np.ascontiguousarray(pk[:, t]), h_out)
                got = h_out
            assert np.array_equal(got, want), (stages, t, np.nonzero((got != want).any(axis=1))[0][:5])
        for ns, agc, vad in hs:
            if ns:
                (O.orc_nsx_release if ns_core else O.orc_ns_release)(ns)
            if agc:
                O.orc_agc_release(agc)
            if vad:
                O.orc_vad_release(vad)
        eng.close()
    # what does not exist at this rate is refused, not approximated
    eng = wmix_b200.Engine(4, 32000, stages=NS)
    assert eng.L.wmixb_set_conferences(eng.h, np.array([0, 4], np.int32).ctypes.data, 1) != 0
    eng.close()
    with pytest.raises(wmix_b200.WmixError):
        wmix_b200.Engine(4, 24000, stages=NS)


def test_nsx_in_the_record_tick_at_wmix_cadence():
    """the daemon's 20 ms record tick (wmixb_record_*) with the switch thrown: NSX on two 10 ms packets -> AGC -> VAD on one
    20 ms packet (the re-staged packet kernel), against the oracle's handles in wmix's own geometry"""
    lib, O = wmix_b200.lib(), oracle()
    O.orc_nsx_init.restype = C.c_void_p
    freq, S, K = 16000, 70, 120
    pkg = freq // 50
    x = make_frames(S, freq, 0, 2 * K, seed=73)
    mic = np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(S, K, pkg)
    play = np.zeros_like(mic)
    stages = NS | AGC | VAD
    eng = wmix_b200.Engine(S, freq, stages=stages, ns_core=1)
    rec = C.c_void_p()
    assert lib.wmixb_record_create(eng.h, 400, C.byref(rec)) == 0
    d_play = torch.zeros((S, pkg), dtype=torch.int16, device=DEV)
    d_mic, d_out = torch.empty_like(d_play), torch.empty_like(d_play)
    d_vad = torch.zeros((S,), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    got = np.empty((K, S, pkg), np.int16)
    for t in range(K):
        d_mic.copy_(torch.from_numpy(np.ascontiguousarray(mic[:, t])))
        assert lib.wmixb_record_tick_device(rec, d_play.data_ptr(), d_mic.data_ptr(), d_out.data_ptr(), d_vad.data_ptr(), None, stages, st) == 0
        got[t] = d_out.cpu().numpy()
    lib.wmixb_record_destroy(rec)
    eng.close()
    for s in list(range(0, S, 3)):
        ns = C.c_void_p(O.orc_nsx_init(1, freq))
        agc = C.c_void_p(O.orc_agc_init(1, freq, 20, 5))
        vad = C.c_void_p(O.orc_vad_init(1, freq, 20))
        for t in range(K):
            f = mic[s, t].copy()
            O.orc_nsx_process(ns, P(f), P(f), pkg)
            assert O.orc_agc_process(agc, P(f), P(f), pkg) == 0
            O.orc_vad_process(vad, P(f), pkg)
            assert np.array_equal(got[t, s], f), (s, t)
        O.orc_nsx_release(ns)
        O.orc_agc_release(agc)
        O.orc_vad_release(vad)
