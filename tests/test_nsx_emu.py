"""CPU check of the fixed-point suppressor's KERNEL BODY (wmix_b200/csrc/nsx.cuh compiled by g++ into the lane-loop
emulator of tests/emu: phases as loops over the 32 lanes, warp reductions as loops between phases) against the oracle:
outputs and the complete per-stream state, bit for bit; the host-built constant tables against the reference's
literals; the parallel two-peak search against the reference's scan."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests._emu import emu
from tests._oracle import P, fnv1a64, nsx_quiet_streams, oracle
from wmix_b200.synth import make_frames

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "nsx.json")))


def _libs():
    O, E = oracle(), emu()
    O.orc_nsx_init_policy.restype = C.c_void_p
    E.emu_nsx_create.restype = C.c_void_p
    E.emu_nsx_record.restype = C.POINTER(C.c_uint32)
    return O, E


def _run(freq, policy, pcm, state_every=1):
    O, E = _libs()
    T, S, n = pcm.shape
    so = np.zeros(4096, np.int32)
    se = np.zeros(4096, np.int32)
    for s in range(S):
        o = C.c_void_p(O.orc_nsx_init_policy(1, freq, policy))
        e = C.c_void_p(E.emu_nsx_create(freq, policy))
        for t in range(T):
            x = pcm[t, s].copy()
            a = np.zeros(n, np.int16)
            b = np.zeros(n, np.int16)
            O.orc_nsx_process(o, P(x), P(a), n)
            E.emu_nsx_frame(e, P(x), P(b))
            assert np.array_equal(a, b), (freq, policy, s, t)
            if t % state_every == 0 or t == T - 1:
                nw = O.orc_nsx_state(o, P(so), 4096)
                assert E.emu_nsx_canonical(freq, E.emu_nsx_record(e), P(se)) == nw
                assert np.array_equal(so[:nw], se[:nw]), (freq, policy, s, t, np.nonzero(so[:nw] != se[:nw])[0][:8])
        O.orc_nsx_release(o)
        E.emu_nsx_destroy(e)


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_kernel_body_vs_oracle_with_state(freq):
    _run(freq, 2, make_frames(6, freq, 0, 1100, seed=9), state_every=7)


@pytest.mark.parametrize("freq", [16000, 8000])
@pytest.mark.parametrize("policy", [0, 1, 3])
def test_nsx_kernel_body_other_policies(freq, policy):
    _run(freq, policy, make_frames(3, freq, 0, 560, seed=10 + policy), state_every=50)


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_kernel_body_quiet_gapped_saturated(freq):
    _run(freq, 2, nsx_quiet_streams(freq, T=460), state_every=5)


def test_nsx_two_peak_search_equals_the_reference_scan():
    _, E = _libs()
    rng = np.random.default_rng(1)

    def scan(h):
        m1 = m2 = p1 = p2 = w1 = w2 = 0
        for i, v in enumerate(h):
            if v > m1:
                m2, w2, p2 = m1, w1, p1
                m1, w1, p1 = v, v, 2 * i + 1
            elif v > m2:
                m2, w2, p2 = v, v, 2 * i + 1
        return [p1, p2, w1, w2]

    for it in range(1500):
        h = np.zeros(1000, np.int16)
        kind = it % 5
        if kind == 0:
            h[rng.integers(0, 1000, size=rng.integers(0, 6))] = rng.integers(0, 4)
        elif kind == 1:
            h[:] = rng.integers(0, 3, size=1000)
        elif kind == 2:
            np.add.at(h, rng.integers(0, 1000, size=512), 1)
        elif kind == 3:
            c = rng.integers(0, 990)
            h[c:c + 8] = rng.integers(0, 200, size=8)
        else:
            h[rng.integers(0, 1000, size=40)] = rng.integers(1, 5, size=40)
        out = np.zeros(4, np.int32)
        E.emu_nsx_two_peaks(P(h), P(out))
        assert list(out) == scan(h.tolist())


def test_nsx_host_tables_match_the_reference_literals():
    """host::nsx_tables (what the engine uploads) against hashes of the literal tables in the reference's sources"""
    _, E = _libs()
    raw = np.zeros(8192, np.uint8)
    for freq, win, n_win in ((16000, "kBlocks160w256x", 256), (8000, "kBlocks80w128x", 128)):
        for policy in (0, 1, 2, 3):
            n = E.emu_nsx_tables(freq, policy, P(raw), raw.size)
            assert n > 0
            tw = raw[:512].view(np.uint32)
            i16 = raw[512:n].view(np.int16)
            window, log_frac, cdiv, log_index = i16[:256], i16[256:512], i16[512:714], i16[714:844]
            f1, f2, sig, lstage = i16[844:1102], i16[1102:1360], i16[1360:1378], i16[1378:1388]
            h = lambda a: fnv1a64(np.ascontiguousarray(a, np.int16).tobytes())
            assert h(window[:n_win]) == G["tables"][win]["hash"]
            assert h(log_frac) == G["tables"]["WebRtcNsx_kLogTableFrac"]["hash"]
            assert h(cdiv[:201]) == G["tables"]["WebRtcNsx_kCounterDiv"]["hash"]
            assert h(log_index[:129]) == G["tables"]["kLogIndex"]["hash"]
            assert h(f1[:257]) == G["tables"]["kFactor1Table"]["hash"]
            if policy:
                assert h(f2[:257]) == G["tables"]["kFactor2Aggressiveness%d" % policy]["hash"]
            assert h(sig[:17]) == G["tables"]["kIndicatorTable"]["hash"]
            assert h(lstage[:9]) == G["tables"]["WebRtcNsx_kLogTable"]["hash"]
            # twiddles = kSinTable1024 sampled at m * 1024 / ana: rebuild the 512 entries the transform can reach from the
            # oracle's table dump (itself pinned to the literal table's hash in test_nsx_oracle_pin.py)
            from tests.test_nsx_oracle_pin import oracle_tables
            sine = oracle_tables()["kSinTable1024"].astype(np.int64)
            step = 1024 // n_win
            m = np.arange(n_win // 2)
            want = ((sine[m * step + 256] & 0xFFFF) << 16) | (sine[m * step] & 0xFFFF)
            assert np.array_equal(tw[:n_win // 2].astype(np.int64), want)


@pytest.mark.parametrize("freq", [16000, 8000])
def test_nsx_kernel_body_second_band(freq):
    """frame<ANA, true> + second_band (wmix's stereo: the right channel as WebRtcNsx's second band) against the oracle's
    two-band handle, zero-input frames included"""
    O, E = _libs()
    n, keep = freq // 100, {16000: 96, 8000: 48}[freq]
    lo = make_frames(3, freq, 0, 600, seed=5)
    hi = make_frames(3, freq, 0, 600, seed=6)
    lo[100:130] = 0
    for s in range(3):
        o = C.c_void_p(O.orc_nsx_init_policy(2, freq, 2))
        e = C.c_void_p(E.emu_nsx_create(freq, 2))
        hb = np.zeros(keep, np.int16)
        for t in range(600):
            x = np.stack([lo[t, s], hi[t, s]], axis=1).reshape(-1).copy()
            y = np.zeros(2 * n, np.int16)
            O.orc_nsx_process(o, P(x), P(y), n)
            a = np.zeros(n, np.int16)
            b = np.zeros(n, np.int16)
            E.emu_nsx_frame_hb(e, P(lo[t, s].copy()), P(a), P(hb), P(hi[t, s].copy()), P(b))
            assert np.array_equal(a, y[0::2]) and np.array_equal(b, y[1::2]), (freq, s, t)
        O.orc_nsx_release(o)
        E.emu_nsx_destroy(e)
