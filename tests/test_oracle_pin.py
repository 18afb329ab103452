"""Pins oracle/liboracle.so (our C restatement) to
  (1) the known-answer vectors transcribed from the reference's own unit tests,
  (2) the unmodified reference compiled here (oracle/_ref/libwmix_ref.so) on seeded inputs,
  (3) the committed fixtures in tests/golden/ (made by tests/golden/make_golden.py from (2)).
CPU only."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests._oracle import P, RefChain, fnv1a64, oracle, ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from wmix_b200.synth import make_frames

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
need_ref = pytest.mark.skipif(ref() is None, reason="oracle/_ref not built (no /root/reference here)")


def i16(*a):
    return np.array(a, dtype=np.int16)


# ---------------------------------------------------------------- G.711
def _g711_all(L, prefix):
    x = np.arange(-32768, 32768, dtype=np.int16)
    ea = np.zeros(65536, np.uint8)
    eu = np.zeros(65536, np.uint8)
    c = np.arange(256, dtype=np.uint8)
    da = np.zeros(256, np.int16)
    du = np.zeros(256, np.int16)
    if prefix:
        assert L.orc_PCM2G711a(P(x), P(ea), 131072) == 65536
        assert L.orc_PCM2G711u(P(x), P(eu), 131072) == 65536
        assert L.orc_G711a2PCM(P(c), P(da), 256) == 512
        assert L.orc_G711u2PCM(P(c), P(du), 256) == 512
    else:
        assert L.PCM2G711a(P(x), P(ea), 131072, 0) == 65536
        assert L.PCM2G711u(P(x), P(eu), 131072, 0) == 65536
        assert L.G711a2PCM(P(c), P(da), 256, 0) == 512
        assert L.G711u2PCM(P(c), P(du), 256, 0) == 512
    return ea, eu, da, du


def test_g711_spot_values_from_survey():
    L = oracle()
    # SURVEY.md §8c spot values (reference probe)
    assert [L.orc_linear2alaw(v) for v in (0, -1, -32768, 32767, -8, 255, 256)] == [0xD5, 0x5A, 0x2A, 0xAA, 0x55, 0xDA, 0xC5]
    assert [L.orc_linear2ulaw(v) for v in (0, -1, -32768, 32767)] == [0xFF, 0x7F, 0x00, 0x80]
    assert L.orc_alaw2linear(0x2A) == -32256 and L.orc_ulaw2linear(0x00) == -32124


def test_g711_golden_hashes():
    g = json.load(open(os.path.join(GOLDEN, "hashes.json")))
    ea, eu, da, du = _g711_all(oracle(), "orc_")
    assert fnv1a64(ea.tobytes()) == g["g711_alaw_enc"]
    assert fnv1a64(eu.tobytes()) == g["g711_ulaw_enc"]
    assert fnv1a64(da.tobytes()) == g["g711_alaw_dec"]
    assert fnv1a64(du.tobytes()) == g["g711_ulaw_dec"]


@need_ref
def test_g711_full_domain_vs_reference():
    for a, b in zip(_g711_all(oracle(), "orc_"), _g711_all(ref(), "")):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------- mix
def test_mix_kat_from_survey():
    L = oracle()
    ring = i16(15648, -25396)
    src = i16(-5670, -4786)
    assert L.orc_mix_same_format(P(ring), 2, 0, P(src), 2, 3) == 0
    assert ring.tolist() == [13758, -26991]
    assert L.orc_volume_add(30000, 30000) == 32767 and L.orc_volume_add(-30000, -30000) == -32768


@need_ref
def test_mix_vs_reference_wmix_load_data():
    R, L = ref(), oracle()
    rng = np.random.default_rng(7)
    ring_bytes = R.oracle_ref_wmix_buff_size()
    n = ring_bytes // 2

    class WPoint(C.Union):
        _fields_ = [("U8", C.c_void_p)]

    R.wmix_load_data.restype = WPoint
    R.wmix_load_data.argtypes = [C.c_void_p, WPoint, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, WPoint,
                                 C.c_uint8, C.POINTER(C.c_uint32)]
    wm = (C.c_uint8 * R.oracle_ref_sizeof_wmix())()
    for rdce_mode, reduce in ((1, 1), (3, 0), (3, 3), (16, 2)):
        ring_ref = (rng.integers(-32768, 32768, n)).astype(np.int16)
        ring_ref[::7] = 0
        ring_orc = ring_ref.copy()
        head_off = (n - 100) * 2
        R.oracle_ref_wmix_seat(wm, P(ring_ref), ring_bytes, rdce_mode, 0, 0)
        src = rng.integers(-32768, 32768, 480).astype(np.int16)
        src[::5] = 0
        tick = C.c_uint32(0)
        head = WPoint(ring_ref.ctypes.data + head_off)
        out = R.wmix_load_data(wm, WPoint(src.ctypes.data), src.nbytes, R.oracle_ref_wmix_freq(), 1, 16, head,
                               reduce, C.byref(tick))
        d = 1 if reduce == rdce_mode else rdce_mode
        pos = L.orc_mix_same_format(P(ring_orc), n, head_off // 2, P(src), len(src), d)
        assert np.array_equal(ring_ref, ring_orc)
        assert out.U8 - ring_ref.ctypes.data == pos * 2


@need_ref
def test_mix_resample_vs_reference_wmix_load_data():
    """different-format branches (R:src/wmix.c:1704-1939) of the real wmix_load_data, mono 16 kHz bus"""
    R, L = ref(), oracle()
    rng = np.random.default_rng(11)
    ring_bytes = R.oracle_ref_wmix_buff_size()
    n = ring_bytes // 2
    mix_freq = R.oracle_ref_wmix_freq()

    class WPoint(C.Union):
        _fields_ = [("U8", C.c_void_p)]

    R.wmix_load_data.restype = WPoint
    R.wmix_load_data.argtypes = [C.c_void_p, WPoint, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, WPoint,
                                 C.c_uint8, C.POINTER(C.c_uint32)]
    L.orc_mix_resample.restype = C.c_uint32
    L.orc_mix_resample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8,
                                   C.c_uint16, C.c_uint8, C.POINTER(C.c_uint32)]
    wm = (C.c_uint8 * R.oracle_ref_sizeof_wmix())()
    cases = [(8000, 1), (8000, 2), (11025, 1), (12000, 2), (16000, 2), (22050, 1), (32000, 2), (44100, 2), (48000, 1),
             (15999, 1), (16001, 2), (300, 1), (65535, 1)]
    for k, (freq, chn) in enumerate(cases):
        rdce_mode, reduce = ((1, 1), (3, 0), (16, 2))[k % 3]
        frames = 331 if freq >= 1000 else 40
        ring_ref = rng.integers(-32768, 32768, n).astype(np.int16)
        ring_ref[::7] = 0
        ring_orc = ring_ref.copy()
        head_off = (n - 150) * 2
        R.oracle_ref_wmix_seat(wm, P(ring_ref), ring_bytes, rdce_mode, 0, 0)
        # the reference reads one frame past the source when it prepares the ramp after the last frame
        src = rng.integers(-32768, 32768, frames * chn + 2).astype(np.int16)
        src[::5] = 0
        if k % 4 == 0:
            src[:] = np.where(rng.random(src.size) < 0.5, 32767, -32768)
        nbytes = frames * chn * 2
        tick = C.c_uint32(0)
        head = WPoint(ring_ref.ctypes.data + head_off)
        out = R.wmix_load_data(wm, WPoint(src.ctypes.data), nbytes, freq, chn, 16, head, reduce, C.byref(tick))
        d = 1 if reduce == rdce_mode else rdce_mode
        wr = C.c_uint32(0)
        pos = L.orc_mix_resample(P(ring_orc), n, head_off // 2, P(src), nbytes, freq, chn, mix_freq, d, C.byref(wr))
        assert np.array_equal(ring_ref, ring_orc), (freq, chn)
        assert out.U8 - ring_ref.ctypes.data == pos * 2, (freq, chn)
        assert tick.value == wr.value * 2, (freq, chn)


@need_ref
def test_play_fifo_vs_reference():
    """playPkgBuff_add / playPkgBuff_get (R:src/wmix.c:482-526): the far-end alignment of the daemon's record tick.
    The reference keeps ONE static ring, so the walk below runs a whole number of ring turns past the start."""
    R, L = ref(), oracle()
    pkg, num = R.oracle_ref_wmix_pkg_size(), R.oracle_ref_wmix_aec_fifo_pkgs()
    delay, interval = R.oracle_ref_wmix_aec_interval_ms(), R.oracle_ref_wmix_interval_ms()
    assert num == delay // interval + 2
    R.playPkgBuff_get.restype = C.c_void_p
    R.playPkgBuff_get.argtypes = [C.c_void_p, C.c_int]
    rng = np.random.default_rng(5)
    fifo = (C.c_uint8 * (16 + 64 * 1280))()
    L.orc_play_fifo_init(fifo, num, pkg)
    zero = np.zeros(pkg, np.uint8)
    for _ in range(num):                      # flush whatever an earlier test left in the static ring
        R.playPkgBuff_add(P(zero))
    for t in range(5 * num):
        x = rng.integers(0, 256, pkg).astype(np.uint8)
        R.playPkgBuff_add(P(x))
        L.orc_play_fifo_add(fifo, P(x))
        for d in (delay, 0, interval, 5 * interval, delay + interval, delay + 5 * interval):
            a, b = np.zeros(pkg, np.uint8), np.zeros(pkg, np.uint8)
            R.playPkgBuff_get(a.ctypes.data, d)
            L.orc_play_fifo_get(fifo, P(b), d // interval)
            assert np.array_equal(a, b), (t, d)
    for _ in range(num):
        R.playPkgBuff_add(P(zero))


@need_ref
def test_wmix_load_data_bookkeeping_vs_reference():
    """the whole wmix_load_data (R:src/wmix.c:1639-1956): head restart for a new / late producer (also past the ring end),
    background-reduce choice, tick arithmetic, the empty 8-bit case — oracle against the real function, a producer's calls
    chained through the returned head and tick"""
    from tests._oracle import MixView, load_data_cases

    R, L = ref(), oracle()
    rng = np.random.default_rng(17)
    ring_bytes, correct, mix_freq = R.oracle_ref_wmix_buff_size(), R.oracle_ref_wmix_play_correct(), R.oracle_ref_wmix_freq()
    n = ring_bytes // 2

    class WPoint(C.Union):
        _fields_ = [("U8", C.c_void_p)]

    R.wmix_load_data.restype = WPoint
    R.wmix_load_data.argtypes = [C.c_void_p, WPoint, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, WPoint, C.c_uint8, C.POINTER(C.c_uint32)]
    L.orc_wmix_load_data.restype = C.c_int32
    L.orc_wmix_load_data.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, C.c_int32, C.c_uint8,
                                     C.POINTER(C.c_uint32)]
    wm = (C.c_uint8 * R.oracle_ref_sizeof_wmix())()
    for play_head, play_tick, reduce_mode in ((1000, 0, 1), (ring_bytes - 200, 5000, 3), (ring_bytes - correct, 777, 16)):
        ring_a = rng.integers(-32768, 32768, n).astype(np.int16)
        ring_b = ring_a.copy()
        R.oracle_ref_wmix_seat(wm, P(ring_a), ring_bytes, reduce_mode, play_head, play_tick)
        view = MixView(ring_bytes, play_head, play_tick, correct, mix_freq, reduce_mode, 1)
        head_a, tick_a = WPoint(None), C.c_uint32(0)
        head_b, tick_b = -1, C.c_uint32(0)
        for k, (freq, chn, sample, frames, reduce) in enumerate(load_data_cases() * 2):
            nbytes = frames * chn * (sample // 8)
            src = rng.integers(-32768, 32768, nbytes // 2 + 4).astype(np.int16)
            if k == 5:                                   # the producer fell behind the play pointer: restart
                tick_a.value = tick_b.value = max(0, play_tick - 1)
            head_a = R.wmix_load_data(wm, WPoint(src.ctypes.data), nbytes, freq, chn, sample, head_a, reduce, C.byref(tick_a))
            head_b = L.orc_wmix_load_data(C.byref(view), P(ring_b), P(src), nbytes, freq, chn, sample, head_b, reduce, C.byref(tick_b))
            off_a = (head_a.U8 - ring_a.ctypes.data) if head_a.U8 else -1
            assert off_a == head_b and tick_a.value == tick_b.value, (play_head, k, off_a, head_b, tick_a.value, tick_b.value)
            assert np.array_equal(ring_a, ring_b), (play_head, k)
        # a stopped mixer or an empty source changes nothing
        view.run = 0
        assert L.orc_wmix_load_data(C.byref(view), P(ring_b), P(src), 64, mix_freq, 1, 16, 40, 0, C.byref(tick_b)) == 40


@need_ref
def test_wmix_struct_prefix_of_include_wmix_h_matches_the_reference_layout(tmp_path):
    """include/wmix.h restates the leading fields of WMix_Struct (R:src/wmixConf.h:176-207) so that wmix_load_data can be
    exported under the reference's own prototype: every field the function reads must sit at the offset the compiled
    reference puts it.  The reference seats a struct with distinctive values; they are read back at OUR offsets."""
    import subprocess

    so = str(tmp_path / "libwmix_layout.so")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "wmix_layout.c"), "-o", so])
    H = C.CDLL(so)
    for f in ("start", "end", "head", "run", "tick", "reduce"):
        getattr(H, "wmixh_off_" + f).restype = C.c_size_t
    H.wmixh_sizeof.restype = C.c_size_t
    R = ref()
    R.oracle_ref_sizeof_wmix.restype = C.c_size_t
    assert H.wmixh_sizeof() <= R.oracle_ref_sizeof_wmix()
    wm = (C.c_uint8 * R.oracle_ref_sizeof_wmix())()
    ring = np.zeros(4000, np.int16)
    R.oracle_ref_wmix_seat(wm, P(ring), 8000, 13, 2468, 0x01020304)
    raw = bytes(wm)

    def u64(off):
        return int.from_bytes(raw[off:off + 8], "little")

    base = ring.ctypes.data
    assert u64(H.wmixh_off_start()) == base
    assert u64(H.wmixh_off_end()) == base + 8000
    assert u64(H.wmixh_off_head()) == base + 2468
    assert raw[H.wmixh_off_run()] == 1
    assert int.from_bytes(raw[H.wmixh_off_tick():H.wmixh_off_tick() + 4], "little") == 0x01020304
    assert raw[H.wmixh_off_reduce()] == 13


# ---------------------------------------------------------------- SPL primitives (reference unit-test KATs)
def test_spl_kats():
    L = oracle()
    # T:.../signal_processing/signal_processing_unittest.cc:92-157
    assert L.orc_norm_w32(111121) == 14 and L.orc_norm_u32(111121) == 15 and L.orc_size_in_bits(111121) == 17
    assert L.orc_norm_w32(0) == 0 and L.orc_norm_w32(-1) == 31 and L.orc_norm_w32(-2147483648) == 0
    assert L.orc_sqrt(1134567892) == 33700
    assert L.orc_div_w32_w16(117, -5) == -23 and L.orc_div_w32_w16(5, 0) == 0x7FFFFFFF
    assert L.orc_sat16(40000) == 32767 and L.orc_sat16(-40000) == -32768


@need_ref
def test_spl_vs_reference_random():
    R, L = ref(), oracle()
    rng = np.random.default_rng(3)
    R.WebRtcSpl_Sqrt.restype = C.c_int32
    for v in list(rng.integers(-2**31, 2**31, 4000)) + [0, 1, 2, 3, 2**31 - 1, -2**31, 73632]:
        assert L.orc_sqrt(int(v)) == R.WebRtcSpl_Sqrt(int(v)), v
    for n in (5, 10, 20, 40, 80):
        for scale in (1, 100, 32767):
            v = (rng.integers(-scale, scale + 1, n)).astype(np.int16)
            if scale == 32767:
                v[0] = -32768
            s1, s2 = C.c_int(0), C.c_int(0)
            assert L.orc_energy(P(v), n, C.byref(s1)) == R.WebRtcSpl_Energy(P(v), n, C.byref(s2))
            assert s1.value == s2.value
    st1 = np.zeros(8, np.int32)
    st2 = np.zeros(8, np.int32)
    for _ in range(50):
        x = rng.integers(-32768, 32768, 8).astype(np.int16)
        o1 = np.zeros(4, np.int16)
        o2 = np.zeros(4, np.int16)
        L.orc_downsample_by2(P(x), 8, P(o1), P(st1))
        R.WebRtcSpl_DownsampleBy2(P(x), 8, P(o2), P(st2))
        assert np.array_equal(o1, o2) and np.array_equal(st1, st2)


# ---------------------------------------------------------------- VAD (reference unit-test KATs)
class VadCore(C.Structure):
    _fields_ = [("ds_state", C.c_int32 * 4), ("noise_means", C.c_int16 * 12), ("speech_means", C.c_int16 * 12),
                ("noise_stds", C.c_int16 * 12), ("speech_stds", C.c_int16 * 12), ("frame_counter", C.c_int32),
                ("over_hang", C.c_int16), ("num_of_speech", C.c_int16), ("age", C.c_int16 * 96),
                ("low_value", C.c_int16 * 96), ("mean_value", C.c_int16 * 6), ("upper_state", C.c_int16 * 5),
                ("lower_state", C.c_int16 * 5), ("hp_state", C.c_int16 * 4), ("oh1", C.c_int16 * 3),
                ("oh2", C.c_int16 * 3), ("individual", C.c_int16 * 3), ("total", C.c_int16 * 3), ("vad", C.c_int)]


def _speech_ii(n):
    return (np.arange(n, dtype=np.int64) ** 2).astype(np.int16)  # (int16_t)(i*i)


def test_vad_filterbank_kat():
    # T:.../vad/vad_filterbank_unittest.cc:26-90
    L = oracle()
    ref_energy = {80: 48, 160: 11, 240: 11}
    ref_feat = {80: [1213, 759, 587, 462, 434, 272], 160: [1479, 1385, 1291, 1200, 1103, 1099],
                240: [1732, 1692, 1681, 1629, 1436, 1436]}
    speech = _speech_ii(240)
    core = VadCore()
    L.orc_vad_core_init(C.byref(core), 0)  # one init, state carried across the three lengths
    for n in (80, 160, 240):
        feat = np.zeros(6, np.int16)
        assert L.orc_vad_features(C.byref(core), P(speech), n, P(feat)) == ref_energy[n]
        assert feat.tolist() == ref_feat[n]
    core = VadCore()
    L.orc_vad_core_init(C.byref(core), 0)
    for n in (80, 160, 240):
        feat = np.zeros(6, np.int16)
        x = np.zeros(240, np.int16)
        assert L.orc_vad_features(C.byref(core), P(x), n, P(feat)) == 0
        assert feat.tolist() == [368, 368, 272, 176, 176, 176]
    for n in (80, 160, 240):
        core = VadCore()
        L.orc_vad_core_init(C.byref(core), 0)
        feat = np.zeros(6, np.int16)
        x = np.ones(240, np.int16)
        assert L.orc_vad_features(C.byref(core), P(x), n, P(feat)) == 0
        assert feat.tolist() == [368, 368, 272, 176, 176, 176]


def test_vad_gmm_kat():
    # T:.../vad/vad_gmm_unittest.cc:21-42
    L = oracle()
    d = C.c_int16(0)
    for (x, m, s), (p, dd) in {(0, 0, 128): (1048576, 0), (16, 128, 128): (1048576, 0), (-16, -128, 128): (1048576, 0),
                               (59, 0, 128): (1024, 7552), (75, 128, 128): (1024, 7552), (-75, -128, 128): (1024, -7552),
                               (105, 0, 128): (0, 13440)}.items():
        assert L.orc_vad_gaussian(x, m, s, C.byref(d)) == p and d.value == dd


def test_vad_sp_kat():
    # T:.../vad/vad_sp_unittest.cc:24-73
    L = oracle()
    zeros = np.zeros(960, np.int16)
    data = _speech_ii(960)
    out = np.zeros(480, np.int16)
    st = np.zeros(2, np.int32)
    L.orc_vad_downsample(P(zeros), P(out), P(st), 960)
    assert st.tolist() == [0, 0] and not out.any()
    L.orc_vad_downsample(P(data), P(out), P(st), 960)
    assert st.tolist() == [207, 2270]
    ref_min = [1600, 720, 509, 512, 532, 552, 570, 588, 606, 624, 642, 659, 675, 691, 707, 723, 1600, 544, 502, 522,
               542, 561, 579, 597, 615, 633, 651, 667, 683, 699, 715, 731]
    core = VadCore()
    L.orc_vad_core_init(C.byref(core), 0)
    for i in range(16):
        v = 500 * (i + 1)
        for ch in range(6):
            assert L.orc_vad_find_minimum(C.byref(core), v, ch) == ref_min[i]
            assert L.orc_vad_find_minimum(C.byref(core), 12000, ch) == ref_min[i + 16]
        core.frame_counter += 1


def test_vad_core_kat():
    # T:.../vad/vad_core_unittest.cc:57-104 — ONE InitCore, then zeros -> 0 and (i*i) -> 1 for every
    # valid (rate, length) pair in kFrameLengths order, state carried from call to call.  The 48 kHz
    # calls of the original are left out (that resampler is not on wmix's path and has its own state).
    L = oracle()
    lengths = [80, 120, 160, 240, 320, 480, 640, 960, 1440]
    speech = _speech_ii(1440)
    zeros = np.zeros(1440, np.int16)
    valid = lambda fs, n: n in (fs // 100, fs // 50, fs * 3 // 100)
    core = VadCore()
    L.orc_vad_core_init(C.byref(core), 0)
    for x, want in ((zeros, 0), (speech, 1)):
        for n in lengths:
            for fs in (8000, 16000, 32000):
                if valid(fs, n):
                    assert L.orc_vad_core_process(C.byref(core), fs, P(x), n) == want, (fs, n, want)
    assert L.orc_vad_core_process(C.byref(core), 9999, P(zeros), 160) == -1
    assert L.orc_vad_core_process(C.byref(core), 16000, P(zeros), 161) == -1
    assert L.orc_vad_core_process(C.byref(core), 16000, None, 160) == -1


# ---------------------------------------------------------------- AGC
def test_agc_gain_table_kat():
    # CalculateGainTable(comp=5,target=0,limiter=0,analogTarget=6)
    L = oracle()
    tab = np.zeros(32, np.int32)
    assert L.orc_agc_analog_target(5) == 6
    assert L.orc_agc_gain_table(P(tab), 5, 0, 0, 6) == 0
    assert tab[:4].tolist() == [74652, 91180, 113772, 126764] and tab[28:].tolist() == [130796] * 4


@need_ref
def test_agc_gain_table_all_gains_vs_reference():
    R, L = ref(), oracle()
    for comp in range(0, 91):
        for lim in (0, 1):
            at = L.orc_agc_analog_target(comp)
            a = np.zeros(32, np.int32)
            b = np.zeros(32, np.int32)
            ra = L.orc_agc_gain_table(P(a), comp, 0, lim, at)
            rb = R.WebRtcAgc_CalculateGainTable(P(b), comp, 0, lim, at)
            assert ra == rb and np.array_equal(a, b), (comp, lim)


# ---------------------------------------------------------------- streams vs the reference
def _streams(freq, n_streams, n_ticks, seed):
    return make_frames(n_streams, freq, 0, n_ticks, seed=seed)  # [T, S, L]


@need_ref
@pytest.mark.parametrize("freq", [8000, 16000, 32000])
@pytest.mark.parametrize("stage", ["vad", "agc", "ns", "chain"])
def test_stage_vs_reference(freq, stage):
    """32000: the handle API's 32 kHz quirks — NS touches only the first 160 samples of each 320-sample packet and
    leaves zeros behind (R:src/webrtc.c:633), AGC runs 5 ms packets (R:src/webrtc.c:727), VAD decimates twice"""
    R, L = ref(), oracle()
    S, T = (6, 700) if stage in ("ns", "chain") else (8, 400)
    if freq == 32000:
        S, T = 3, 300
    x = _streams(freq, S, T, seed=11)
    kw = dict(ns=stage in ("ns", "chain"), agc=stage in ("agc", "chain"), vad=stage in ("vad", "chain"))
    for s in range(S):
        pcm = np.ascontiguousarray(x[:, s, :]).reshape(-1)
        a = RefChain(R, freq, **kw)
        b = RefChain(L, freq, prefix="orc_", **kw)
        ya, yb = a.run(pcm), b.run(pcm)
        a.close()
        b.close()
        assert np.array_equal(ya, yb), (stage, freq, s, int(np.abs(ya.astype(int) - yb).max()))


@need_ref
@pytest.mark.parametrize("freq", [8000, 16000, 32000])
def test_ns_stereo_right_channel_as_high_band_vs_reference(freq):
    """ns_init(2, ..): wmix hands the right channel to WebRtcNs as a second band (R:src/webrtc.c:624-636), which only
    gets the time-domain high-band gain (T:.../ns/ns_core.c:1361-1414) and the analysis-buffer delay"""
    R, L = ref(), oracle()
    n = freq // 100
    T = 320
    x = _streams(min(freq, 16000), 4, T * (2 if freq == 32000 else 1), seed=19)
    pcm = np.ascontiguousarray(x.transpose(1, 0, 2)).reshape(4, -1)       # 4 mono streams
    for a, b in ((0, 1), (2, 3), (3, 3)):
        st = np.empty(2 * pcm.shape[1], np.int16)
        st[0::2], st[1::2] = pcm[a], pcm[b]
        hr = C.c_void_p(R.ns_init(2, freq, None))
        ho = C.c_void_p(L.orc_ns_init(2, freq))
        assert hr and ho
        for t in range(len(st) // (2 * n)):
            f = st[t * 2 * n:(t + 1) * 2 * n]
            ya, yb = np.zeros(2 * n, np.int16), np.zeros(2 * n, np.int16)
            R.ns_process(hr, P(f.copy()), P(ya), n)
            L.orc_ns_process(ho, P(f.copy()), P(yb), n)
            assert np.array_equal(ya, yb), (freq, a, b, t)
        R.ns_release(hr)
        L.orc_ns_release(ho)


@need_ref
def test_config1_wav_vs_reference():
    """BASELINE config 1: NS on audio/1x8000.wav through src/webrtc.c, then AGC(5) and VAD(10 ms) in place."""
    wav = "/root/reference/audio/1x8000.wav"
    if not os.path.exists(wav):
        pytest.skip("fixture wav not on this box")
    pcm = np.fromfile(wav, dtype=np.int16, offset=44)
    pcm = pcm[: len(pcm) // 80 * 80]
    a = RefChain(ref(), 8000)
    b = RefChain(oracle(), 8000, prefix="orc_")
    ya, yb = a.run(pcm), b.run(pcm)
    assert np.array_equal(ya, yb)
    g = json.load(open(os.path.join(GOLDEN, "hashes.json")))
    assert fnv1a64(yb.tobytes()) == g["config1_ns_agc_vad"]


def test_golden_streams():
    """Committed fixtures: first/last samples + hash of reference output on seeded streams."""
    g = json.load(open(os.path.join(GOLDEN, "hashes.json")))
    for key, spec in g["streams"].items():
        freq, stage, S, T, seed = spec["freq"], spec["stage"], spec["n_streams"], spec["n_ticks"], spec["seed"]
        x = _streams(freq, S, T, seed)
        kw = dict(ns=stage in ("ns", "chain"), agc=stage in ("agc", "chain"), vad=stage in ("vad", "chain"))
        outs = []
        for s in range(S):
            c = RefChain(oracle(), freq, prefix="orc_", **kw)
            outs.append(c.run(np.ascontiguousarray(x[:, s, :]).reshape(-1)))
            c.close()
        y = np.stack(outs)
        assert fnv1a64(y.tobytes()) == spec["hash"], key


# ---------------------------------------------------------------- AEC
def test_golden_handle_quirks_mix_resample_and_play_fifo():
    """Committed fixtures made from the unmodified reference (tests/golden/make_golden.py) for the paths added after the first
    capture: 32 kHz handles, stereo NS, the resampling branches of the real wmix_load_data, playPkgBuff_get.  Needs no
    reference at run time."""
    L = oracle()
    g = json.load(open(os.path.join(GOLDEN, "hashes.json")))
    for key, spec in g["handles"].items():
        T, seed = spec["n_ticks"], spec["seed"]
        if key.startswith("ns_stereo"):
            freq = spec["freq"]
            n = freq // 100
            xf = _streams(freq, 4, T, seed)
            outs = []
            for a, b in spec["pairs"]:
                h = C.c_void_p(L.orc_ns_init(2, freq))
                st = np.empty((T, 2 * n), np.int16)
                st[:, 0::2], st[:, 1::2] = xf[:, a], xf[:, b]
                out = np.zeros_like(st)
                for t in range(T):
                    L.orc_ns_process(h, P(st[t].copy()), P(out[t]), n)
                L.orc_ns_release(h)
                outs.append(out.reshape(-1))
            y = np.stack(outs)
        else:
            S, stage = spec["n_streams"], spec["stage"]
            xs = _streams(16000, 2 * S, 2 * T, seed)
            kw = dict(ns=stage in ("ns", "chain"), agc=stage in ("agc", "chain"), vad=stage in ("vad", "chain"))
            outs = []
            for s_ in range(S):
                c = RefChain(L, 32000, prefix="orc_", **kw)
                outs.append(c.run(np.ascontiguousarray(xs[:, s_, :]).reshape(-1)))
                c.close()
            y = np.stack(outs)
        assert fnv1a64(y.tobytes()) == spec["hash"], key
        assert y[:, -8:].tolist() == spec["tail"], key
    # resample-on-mix
    m = g["mix_resample"]
    L.orc_mix_resample.restype = C.c_uint32
    L.orc_mix_resample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint16, C.c_uint8,
                                   C.c_uint16, C.c_uint8, C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(m["seed"])
    n = m["ring_samples"]
    for cs in m["cases"]:
        ring = rng.integers(-32768, 32768, n).astype(np.int16)
        src = rng.integers(-32768, 32768, cs["frames"] * cs["chn"] + 2).astype(np.int16)
        assert fnv1a64(ring.tobytes()) == cs["ring_in"] and fnv1a64(src.tobytes()) == cs["src_in"]
        wr = C.c_uint32(0)
        pos = L.orc_mix_resample(P(ring), n, cs["head"], P(src), cs["frames"] * cs["chn"] * 2, cs["freq"], cs["chn"], m["mix_freq"],
                                 cs["rdce"], C.byref(wr))
        assert fnv1a64(ring.tobytes()) == cs["ring_out"] and pos == cs["new_head"] and wr.value == cs["written"], cs
    # play FIFO: which add does playPkgBuff_get(AEC_INTERVALMS) return after each add
    f = g["play_fifo"]
    fifo = (C.c_uint8 * (16 + 64 * 1280))()
    L.orc_play_fifo_init(fifo, f["n_pkg"], 8)
    for t, want in enumerate(f["got_tag"]):
        L.orc_play_fifo_add(fifo, P(np.full(8, (t % 250) + 1, np.uint8)))
        b = np.zeros(8, np.uint8)
        L.orc_play_fifo_get(fifo, P(b), f["delay_pkgs"])
        assert int(b[0]) == want, t


@need_ref
def test_aec_tables_vs_reference_symbols():
    """The oracle builds the AEC tables from formulas (+ eight one-ulp corrections of rdft_w); the reference
    exports its literals (T:.../aec/aec_rdft.c:32-49, aec_core.c:49-96) — they must be identical."""
    R, L = ref(), oracle()
    w = np.zeros(64, np.float32)
    h = np.zeros(65, np.float32)
    wc = np.zeros(65, np.float32)
    od = np.zeros(65, np.float32)
    L.orc_aec_tables(P(w), P(h), P(wc), P(od))

    def sym(name, n):
        return np.array((C.c_float * n).in_dll(R, name), dtype=np.float32)

    assert np.array_equal(w, sym("rdft_w", 64))
    assert np.array_equal(h, sym("WebRtcAec_sqrtHanning", 65))
    assert np.array_equal(wc, sym("WebRtcAec_weightCurve", 65))
    assert np.array_equal(od, sym("WebRtcAec_overDriveCurve", 65))


@need_ref
def test_aec_rdft_vs_reference():
    R, L = ref(), oracle()
    R.aec_rdft_init()
    rng = np.random.default_rng(5)
    for k in range(300):
        a = (rng.standard_normal(128) * 10 ** rng.uniform(-3, 4)).astype(np.float32)
        for inv in (0, 1):
            x, y = a.copy(), a.copy()
            L.orc_aec_rdft(P(x), inv)
            (R.aec_rdft_inverse_128 if inv else R.aec_rdft_forward_128)(P(y))
            assert np.array_equal(x.view(np.int32), y.view(np.int32))


@need_ref
@pytest.mark.parametrize("freq,ims,T,delay", [(8000, 10, 1200, 0), (8000, 20, 500, 0), (16000, 10, 600, 0),
                                              (8000, 10, 700, 60), (16000, 10, 500, 240)])
def test_aec_vs_reference_pairs(freq, ims, T, delay):
    """aec_process2 on echo + local talker + noise pairs (config 4's signal model), bit for bit; the streams
    include the zero-far-end and full-scale cohorts."""
    from tests._oracle import aec_run_pairs
    from wmix_b200.synth import make_aec_pairs

    S = 5
    far, near = make_aec_pairs(S, freq, 0, T, seed=17)
    n = freq // 1000 * (20 if (freq == 8000 and ims == 20) else 10)
    far = far.transpose(1, 0, 2).reshape(S, -1, n).transpose(1, 0, 2)
    near = near.transpose(1, 0, 2).reshape(S, -1, n).transpose(1, 0, 2)
    a = aec_run_pairs(ref(), "", far, near, freq, ims, delay)
    b = aec_run_pairs(oracle(), "orc_", far, near, freq, ims, delay)
    assert np.array_equal(a, b)
    # the canceller really cancels: the residual of the echo-only tail is far below the near-end level
    assert np.abs(a[-100:].astype(float)).mean() < 0.5 * np.abs(near[-100:].astype(float)).mean()


@need_ref
def test_aec_vs_reference_irregular_cadence_and_errors():
    """aec_setFrameFar / aec_process in bursts (exercises the stuffing / flush paths of the far ring), a
    delay that changes mid-call, and the wrapper's error returns."""
    from tests._oracle import AecRef
    from wmix_b200.synth import make_aec_pairs

    R, L = ref(), oracle()
    for freq in (8000, 16000):
        n = freq // 100
        far, near = make_aec_pairs(2, freq, 0, 600, seed=23)
        for s in range(2):
            a, b = AecRef(R, freq), AecRef(L, freq, prefix="orc_")
            t = 0
            while t < 600:
                grp = min(600 - t, 1 + (t * 7 + s) % 4)
                for g in range(grp):
                    assert a.set_far(far[t + g, s]) == b.set_far(far[t + g, s]) == 0
                for g in range(grp):
                    d = 0 if t < 200 else 100
                    ya, ra = a.process(near[t + g, s], d)
                    yb, rb = b.process(near[t + g, s], d)
                    assert ra == rb == 0 and np.array_equal(ya, yb), (freq, s, t)
                t += grp
            # out-of-range delay: processed, then reported as an error, output untouched (R:src/webrtc.c:382-387)
            ya, ra = a.process2(far[0, s], near[0, s], 600)
            yb, rb = b.process2(far[0, s], near[0, s], 600)
            assert ra == rb == -1 and not ya.any() and not yb.any()
            ya, ra = a.process2(far[1, s], near[1, s], 0)
            yb, rb = b.process2(far[1, s], near[1, s], 0)
            assert ra == rb == 0 and np.array_equal(ya, yb)
            a.close()
            b.close()
    assert not L.orc_aec_init(1, 32000, 10) and not R.aec_init(1, 32000, 10, None)


def test_golden_aec_streams():
    """Committed fixtures made from the reference's aec_process2 (tests/golden/make_golden.py)."""
    from tests._oracle import aec_run_pairs
    from wmix_b200.synth import make_aec_pairs

    g = json.load(open(os.path.join(GOLDEN, "hashes.json")))
    for key, spec in g["aec"].items():
        far, near = make_aec_pairs(spec["n_streams"], spec["freq"], 0, spec["n_ticks"], seed=spec["seed"])
        y = aec_run_pairs(oracle(), "orc_", far, near, spec["freq"], 10, spec["delay_ms"])
        y = np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(spec["n_streams"], -1)
        assert fnv1a64(y.tobytes()) == spec["hash"], key
        assert y[:, -8:].tolist() == spec["tail"]


# ---------------------------------------------------------------- wmix_pcm_zoom / wmix_len_of_* (R:src/wmix.c:49-222)
ZOOM_CASES = [(1, 16000, 1, 8000), (1, 8000, 1, 16000), (2, 16000, 1, 8000), (1, 8000, 2, 16000), (2, 44100, 1, 8000),
              (1, 8000, 1, 44100), (1, 22050, 2, 16000), (2, 32000, 2, 16000), (2, 8000, 2, 48000), (1, 16000, 1, 16000),
              (2, 16000, 2, 16000), (1, 11025, 1, 8000), (1, 8000, 1, 11025), (2, 48000, 1, 16000), (1, 16000, 2, 16000)]


@pytest.mark.parametrize("ic,ifr,oc,ofr", ZOOM_CASES)
def test_zoom_oracle_vs_reference(ic, ifr, oc, ofr):
    """the restated phase walker against the reference's own wmix_pcm_zoom / wmix_len_of_in / wmix_len_of_out"""
    R = ref()
    if R is None:
        pytest.skip("reference not built here")
    L = oracle()
    for f in (R.wmix_len_of_out, R.wmix_len_of_in, R.wmix_pcm_zoom, L.orc_len_of_out, L.orc_len_of_in, L.orc_pcm_zoom):
        f.restype = C.c_uint32
    rng = np.random.default_rng(ic * 1000 + oc + ifr + ofr)
    for in_bytes in (2 * ic, 320 * ic, 640, 1764 * 2, 4000, 2 * ic * 777):
        x = rng.integers(-32768, 32768, in_bytes // 2).astype(np.int16)
        cap = 16 * in_bytes * max(1, ofr // ifr + 1) + 64
        a = np.zeros(cap, np.uint8)
        b = np.zeros(cap, np.uint8)
        na = R.wmix_pcm_zoom(ic, ifr, P(x.copy()), in_bytes, oc, ofr, P(a))
        nb = L.orc_pcm_zoom(ic, ifr, P(x.copy()), in_bytes, oc, ofr, P(b))
        assert na == nb and np.array_equal(a[:na], b[:nb]), (in_bytes, na, nb)
        assert R.wmix_len_of_out(ic, ifr, in_bytes, oc, ofr) == L.orc_len_of_out(ic, ifr, in_bytes, oc, ofr)
        assert R.wmix_len_of_in(ic, ifr, oc, ofr, in_bytes) == L.orc_len_of_in(ic, ifr, oc, ofr, in_bytes)


def test_zoom_known_answers():
    """hand-checkable cases of R:src/wmix.c:139-222: 16k -> 8k mono keeps every second sample starting with the second
    (the accumulator reaches 1.0 on the second step); 8k -> 16k repeats each sample; mono -> stereo duplicates; the
    stereo -> stereo rate change writes nothing (dead 0x22 case)."""
    L = oracle()
    L.orc_pcm_zoom.restype = C.c_uint32
    x = np.arange(1, 17, dtype=np.int16)
    out = np.zeros(64, np.int16)
    n = L.orc_pcm_zoom(1, 16000, P(x), 32, 1, 8000, P(out))
    assert n == 16 and out[:8].tolist() == [2, 4, 6, 8, 10, 12, 14, 16]
    n = L.orc_pcm_zoom(1, 8000, P(x), 8, 1, 16000, P(out))
    assert n == 16 and out[:8].tolist() == [1, 1, 2, 2, 3, 3, 4, 4]
    n = L.orc_pcm_zoom(1, 8000, P(x), 4, 2, 16000, P(out))
    assert n == 16 and out[:8].tolist() == [1, 1, 1, 1, 2, 2, 2, 2]
    n = L.orc_pcm_zoom(2, 16000, P(x), 32, 1, 8000, P(out))
    assert n == 8 and out[:4].tolist() == [3, 7, 11, 15]
    assert L.orc_pcm_zoom(2, 32000, P(x), 32, 2, 16000, P(out)) == 0


# ---------------------------------------------------------------- RTP framing (R:src/rtp.h:51-70, R:src/rtp.c:20-99)
def test_rtp_header_known_answer():
    """RFC 3550 layout as the reference's bit fields produce it: V=2 PCMA marker seq=0x1234 ts=0x01020304 ssrc=0x0A0B0C0D"""
    L = oracle()
    b = np.zeros(12, np.uint8)
    L.orc_rtp_header_bytes(P(b), 0, 0, 0, 2, 8, 1, 0x1234, 0x01020304, 0x0A0B0C0D)
    assert b.tolist() == [0x80, 0x88, 0x12, 0x34, 1, 2, 3, 4, 0x0A, 0x0B, 0x0C, 0x0D]
    L.orc_rtp_header_bytes(P(b), 3, 1, 1, 2, 0, 0, 65535, 0xFFFFFFFF, 1)
    assert b.tolist() == [0xB3, 0x00, 0xFF, 0xFF, 255, 255, 255, 255, 0, 0, 0, 1]


def test_rtp_oracle_vs_reference_over_loopback():
    """the reference's own rtp_header / rtp_send / rtp_recv (R:src/rtp.c) through a UDP socket pair on 127.0.0.1:
    the bytes on the wire, the seq++ / timestamp rule of the PCMA send loop (R:src/wmixTask.c:1139-1143) and what the
    receiver reports must be what the oracle states"""
    import socket

    R = ref()
    if R is None:
        pytest.skip("reference not built here")
    L = oracle()
    R.rtp_socket.restype = C.c_void_p
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0))
    rx.settimeout(5)
    port = rx.getsockname()[1]
    ss = C.c_void_p(R.rtp_socket(b"127.0.0.1", port, False))
    assert ss
    pkt = np.zeros(12 + 4096, np.uint8)                      # RtpPacket
    R.rtp_header(P(pkt), 0, 0, 0, 2, 8, 1, 0, 0, 0)          # R:src/wmixTask.c:1058
    ts, seq = C.c_uint32(0), C.c_uint16(0)
    rng = np.random.default_rng(5)
    want = np.zeros(172, np.uint8)
    for k in range(70000 // 160 + 5):
        codes = rng.integers(0, 256, 160).astype(np.uint8)
        # the reference's loop body: payload, timestamp += ret / chn, rtp_send (which then does seq++)
        pkt[12:172] = codes
        hdr_ts = pkt[4:8].view(np.uint32)
        hdr_ts[0] = hdr_ts[0] + 160
        assert R.rtp_send(ss, P(pkt), 160) == 172
        got = np.frombuffer(rx.recv(4096), np.uint8)
        L.orc_rtp_send_step(C.byref(ts), 0, C.byref(seq), 8, 1, 1, P(codes), 160, P(want))
        assert np.array_equal(got, want), k
        # parse side
        s2, t2, c2, pt, m = C.c_uint16(), C.c_uint32(), C.c_uint32(), C.c_uint8(), C.c_uint8()
        assert L.orc_rtp_parse(P(want), C.byref(s2), C.byref(t2), C.byref(c2), C.byref(pt), C.byref(m)) == 160
        assert (s2.value, t2.value, c2.value, pt.value, m.value) == (k & 0xFFFF, 160 * (k + 1), 0, 8, 1)
    # the reference's receiver: a PCMA datagram of any length is reported as 160 payload bytes (R:src/rtp.c:89-91)
    rs = C.c_void_p(R.rtp_socket(b"127.0.0.1", 0, True))
    R.rtp_socket_close(C.byref(ss))
    rx.close()
