"""CPU pin of the fixed-point suppressor's oracle (oracle/orc_nsx.c): against the compiled reference's WebRtcNsx_* and
its handle layer built with MAKE_WEBRTC_NSX when oracle/_ref exists, and ALWAYS against tests/golden/nsx.json — outputs
and table hashes recorded from the unmodified reference by tests/golden/make_nsx.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests._oracle import (NsxCore, P, fnv1a64, nsx_handle_run, nsx_quiet_streams, oracle, ref, ref_nsx)
from wmix_b200.synth import make_frames

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
G = json.load(open(os.path.join(GOLDEN, "nsx.json")))
TABLE_ORDER = ["WebRtcNsx_kLogTable", "WebRtcNsx_kCounterDiv", "WebRtcNsx_kLogTableFrac", "kBlocks80w128x", "kBlocks160w256x",
               "kFactor1Table", "kFactor2Aggressiveness1", "kFactor2Aggressiveness2", "kFactor2Aggressiveness3", "kSumLogIndex",
               "kSumSquareLogIndex", "kLogIndex", "kDeterminantEstMatrix", "kSinTable1024", "kIndicatorTable"]


def oracle_core_run(freq, pcm, policy=2, hb=None):
    """[T, S, n] through one oracle handle per stream (mono, or two bands when hb is given)"""
    O = oracle()
    O.orc_nsx_init_policy.restype = C.c_void_p
    T, S, n = pcm.shape
    out = np.zeros_like(pcm)
    out_hb = np.zeros_like(pcm) if hb is not None else None
    for s in range(S):
        h = C.c_void_p(O.orc_nsx_init_policy(1 if hb is None else 2, freq, policy))
        for t in range(T):
            if hb is None:
                x = pcm[t, s].copy()
                O.orc_nsx_process(h, P(x), P(out[t, s]), n)
            else:
                x = np.stack([pcm[t, s], hb[t, s]], axis=1).reshape(-1).copy()
                y = np.zeros(2 * n, np.int16)
                O.orc_nsx_process(h, P(x), P(y), n)
                out[t, s], out_hb[t, s] = y[0::2], y[1::2]
        O.orc_nsx_release(h)
    return out if hb is None else (out, out_hb)


def check(desc, y):
    y = np.ascontiguousarray(y)
    assert list(y.shape) == desc["shape"]
    flat = y.reshape(-1)
    assert flat[:8].tolist() == desc["head"] and flat[-8:].tolist() == desc["tail"]
    assert fnv1a64(y.tobytes()) == desc["hash"]


def oracle_tables():
    O = oracle()
    n = O.orc_nsx_tables(None, 0)
    t = np.zeros(n, np.int16)
    assert O.orc_nsx_tables(P(t), n) == n
    out, k = {}, 0
    for name in TABLE_ORDER:
        m = G["tables"][name]["n"]
        out[name] = t[k:k + m].copy()
        k += m
    assert k == n
    return out


def test_nsx_tables_match_the_reference_literals():
    """the closed forms of orc_nsx.c reproduce every literal table of nsx_core.c / complex_fft_tables.h (hashes recorded
    from the reference's sources; entry 0 of the four log-index tables is 'invalid' there and zero in both)"""
    for name, t in oracle_tables().items():
        assert fnv1a64(t.tobytes()) == G["tables"][name]["hash"], name


def test_nsx_tables_vs_reference_sources_when_present():
    if not os.path.exists("/root/reference/pkg/webrtc_cut.tar.gz"):
        pytest.skip("reference sources not on this box")
    from tests.golden.make_nsx import reference_tables
    ours = oracle_tables()
    for name, lit in zip(TABLE_ORDER, reference_tables()):
        assert np.array_equal(ours[name], lit), name


@pytest.mark.parametrize("freq", [8000, 16000])
@pytest.mark.parametrize("policy", [0, 1, 2, 3])
def test_nsx_golden_core(freq, policy):
    d = G["core"]["synth_%d_p%d" % (freq, policy)]
    pcm = make_frames(d["n_streams"], freq, 0, d["n_ticks"], seed=d["seed"])
    check(d, oracle_core_run(freq, pcm, policy))


@pytest.mark.parametrize("freq", [8000, 16000])
def test_nsx_golden_quiet_gapped_saturated(freq):
    check(G["core"]["quiet_%d" % freq], oracle_core_run(freq, nsx_quiet_streams(freq)))


@pytest.mark.parametrize("freq", [8000, 16000])
def test_nsx_golden_second_band(freq):
    d = G["core"]["stereo_%d" % freq]
    lo, hi = oracle_core_run(freq, make_frames(4, freq, 0, 1100, seed=d["seed"]), 2, make_frames(4, freq, 0, 1100, seed=d["seed_hb"]))
    check(d["lo"], lo)
    check(d["hi"], hi)


@pytest.mark.parametrize("freq", [8000, 16000, 32000])
@pytest.mark.parametrize("chn", [1, 2])
def test_nsx_golden_handle_layer(chn, freq):
    d = G["handle"]["%d_%d" % (chn, freq)]
    x = make_frames(chn, freq, 0, d["n_ticks"], seed=d["seed"])
    pcm = np.ascontiguousarray(x.transpose(0, 2, 1).reshape(d["n_ticks"], -1))
    check(d, nsx_handle_run(oracle(), "orc_nsx", chn, freq, pcm))


def test_nsx_golden_config1_wav():
    wav = np.fromfile(os.path.join(GOLDEN, "config1_in_20s.s16"), np.int16).reshape(-1, 80)
    check(G["config1_nsx"], nsx_handle_run(oracle(), "orc_nsx", 1, 8000, wav))


@pytest.mark.parametrize("freq", [8000, 16000])
def test_nsx_oracle_vs_reference_fresh_seed(freq):
    """seeds the fixtures do not hold, straight against WebRtcNsx_* (frame by frame: past frames 50, 200 and the
    512-frame threshold re-learning)"""
    R = ref()
    if R is None:
        pytest.skip("oracle/_ref not built")
    O = oracle()
    O.orc_nsx_init_policy.restype = C.c_void_p
    n = freq // 100
    pcm = make_frames(12, freq, 0, 650, seed=77)
    for s in range(12):
        c = NsxCore(R, freq, 2)
        h = C.c_void_p(O.orc_nsx_init_policy(1, freq, 2))
        for t in range(650):
            y = np.zeros(n, np.int16)
            x = pcm[t, s].copy()
            O.orc_nsx_process(h, P(x), P(y), n)
            assert np.array_equal(c.frame(pcm[t, s]), y), (s, t)
        O.orc_nsx_release(h)
        c.close()


@pytest.mark.parametrize("chn,freq", [(1, 8000), (2, 16000), (1, 32000)])
def test_nsx_handle_layer_vs_reference_built_with_the_switch(chn, freq):
    RX = ref_nsx()
    if RX is None:
        pytest.skip("oracle/_ref not built")
    x = make_frames(chn, freq, 0, 250, seed=31)
    pcm = np.ascontiguousarray(x.transpose(0, 2, 1).reshape(250, -1))
    assert np.array_equal(nsx_handle_run(RX, "", chn, freq, pcm), nsx_handle_run(oracle(), "orc_nsx", chn, freq, pcm))
