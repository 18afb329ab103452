"""ctypes loaders for the two checkers: oracle/liboracle.so (our C restatement) and
oracle/_ref/libwmix_ref.so (the unmodified reference, when it was built in this tree).
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def fnv1a64(b: bytes) -> str:
    h = 0xCBF29CE484222325
    for x in b:
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, so])
        L = C.CDLL(so)
        for f in ("orc_vad_init", "orc_agc_init", "orc_ns_init", "orc_aec_init"):
            getattr(L, f).restype = C.c_void_p
        L.orc_ns_prior_model.restype = C.POINTER(C.c_float)
        L.orc_mix_same_format.restype = C.c_uint32
        for f in ("orc_linear2alaw", "orc_linear2ulaw"):
            getattr(L, f).restype = C.c_uint8
        for f in ("orc_alaw2linear", "orc_ulaw2linear", "orc_norm_w32", "orc_norm_u32", "orc_size_in_bits",
                  "orc_sat16", "orc_vad_find_minimum", "orc_vad_features", "orc_agc_process_vad",
                  "orc_volume_add", "orc_agc_analog_target"):
            getattr(L, f).restype = C.c_int16
        _oracle = L
    return _oracle


def ref():
    """The real reference, or None when oracle/_ref was not built (no /root/reference)."""
    global _ref
    if _ref is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libwmix_ref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        for f in ("ns_init", "vad_init", "agc_init", "aec_init"):
            getattr(L, f).restype = C.c_void_p
        _ref = L
    return _ref


_ref_nsx = None


def ref_nsx():
    """The reference with its own NS switch thrown (R:src/webrtc.c:512, MAKE_WEBRTC_NSX): ns_init / ns_process run the
    fixed-point core.  None when oracle/_ref was not built."""
    global _ref_nsx
    if _ref_nsx is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libwmix_ref_nsx.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        for f in ("ns_init", "vad_init", "agc_init", "aec_init"):
            getattr(L, f).restype = C.c_void_p
        _ref_nsx = L
    return _ref_nsx


class NsxCore:
    """WebRtcNsx_* of the compiled reference driven directly (any policy; one or two bands)."""

    def __init__(self, L, freq, policy=2):
        self.L = L
        self.h = C.c_void_p()
        assert L.WebRtcNsx_Create(C.byref(self.h)) == 0
        assert L.WebRtcNsx_Init(self.h, freq) == 0
        assert L.WebRtcNsx_set_policy(self.h, policy) == 0
        self.n = 80 if freq == 8000 else 160

    def frame(self, lo, hi=None):
        lo = np.ascontiguousarray(lo, np.int16)
        o = np.zeros(self.n, np.int16)
        if hi is None:
            self.L.WebRtcNsx_Process(self.h, (C.c_void_p * 1)(lo.ctypes.data), 1, (C.c_void_p * 1)(o.ctypes.data))
            return o
        hi = np.ascontiguousarray(hi, np.int16)
        oh = np.zeros(self.n, np.int16)
        self.L.WebRtcNsx_Process(self.h, (C.c_void_p * 2)(lo.ctypes.data, hi.ctypes.data), 2,
                                 (C.c_void_p * 2)(o.ctypes.data, oh.ctypes.data))
        return o, oh

    def close(self):
        self.L.WebRtcNsx_Free(self.h)


def nsx_handle_run(L, prefix, chn, freq, pcm, policy=None):
    """pcm int16 [T, n*chn] (interleaved) through ns_init / ns_process of a checker: the reference built with
    MAKE_WEBRTC_NSX (prefix "") or the oracle (prefix "orc_nsx")."""
    pcm = np.ascontiguousarray(pcm, np.int16)
    T = pcm.shape[0]
    n = freq // 100
    out = np.zeros_like(pcm)
    if prefix:
        L.orc_nsx_init.restype = C.c_void_p
        L.orc_nsx_init_policy.restype = C.c_void_p
        h = C.c_void_p(L.orc_nsx_init(chn, freq) if policy is None else L.orc_nsx_init_policy(chn, freq, policy))
        assert h
        for t in range(T):
            L.orc_nsx_process(h, P(pcm[t]), P(out[t]), n)
        L.orc_nsx_release(h)
    else:
        h = C.c_void_p(L.ns_init(chn, freq, None))
        assert h
        for t in range(T):
            L.ns_process(h, P(pcm[t]), P(out[t]), n)
        L.ns_release(h)
    return out


def nsx_quiet_streams(freq, T=700, seed=5):
    """near-silent / gapped / full-scale corner streams [T, 8, n]: amplitudes 1, 2, 3 reach the modulo-32 shift of
    nsx_core.c:1140 / :1716 on x86, stream 6 saturates, all see 30 all-zero frames"""
    rng = np.random.default_rng(seed)
    n = freq // 100
    q = np.zeros((T, 8, n), np.int16)
    for s, amp in enumerate([1, 2, 3, 8, 40, 300, 32767, 5]):
        q[:, s] = rng.integers(-amp, amp + 1, size=(T, n))
    q[100:130] = 0
    q[300:310, :, ::2] = 0
    q[400:420, 6] = 32767
    q[420:440, 6] = -32768
    return q


class RefChain:
    """NS -> AGC -> VAD through the reference's own handle API, one 10 ms frame at a time."""

    def __init__(self, L, freq, ns=True, agc=True, vad=True, gain=5, prefix=""):
        self.L, self.freq, self.n = L, freq, freq // 100
        g = lambda name: getattr(L, prefix + name)
        self.g = g
        self.ns = C.c_void_p(g("ns_init")(1, freq, None) if prefix == "" else g("ns_init")(1, freq)) if ns else None
        if prefix == "":
            self.agc = C.c_void_p(g("agc_init")(1, freq, 10, gain, None)) if agc else None
            self.vad = C.c_void_p(g("vad_init")(1, freq, 10, None)) if vad else None
        else:
            self.agc = C.c_void_p(g("agc_init")(1, freq, 10, gain)) if agc else None
            self.vad = C.c_void_p(g("vad_init")(1, freq, 10)) if vad else None

    def frame(self, x):
        x = np.ascontiguousarray(x, dtype=np.int16).copy()
        if self.ns:
            self.g("ns_process")(self.ns, P(x), P(x), self.n)
        if self.agc:
            self.g("agc_process")(self.agc, P(x), P(x), self.n)
        if self.vad:
            self.g("vad_process")(self.vad, P(x), self.n)
        return x

    def run(self, pcm):
        pcm = np.asarray(pcm, dtype=np.int16)
        nf = len(pcm) // self.n
        out = np.empty(nf * self.n, dtype=np.int16)
        for i in range(nf):
            out[i * self.n:(i + 1) * self.n] = self.frame(pcm[i * self.n:(i + 1) * self.n])
        return out

    def close(self):
        if self.ns:
            self.g("ns_release")(self.ns)
        if self.agc:
            self.g("agc_release")(self.agc)
        if self.vad:
            self.g("vad_release")(self.vad)


class AecRef:
    """aec_process2 / aec_setFrameFar / aec_process through a checker (oracle: prefix "orc_", reference: "")."""

    def __init__(self, L, freq, interval_ms=10, prefix=""):
        self.L, self.prefix = L, prefix
        self.pkg = freq // 1000 * (20 if (freq <= 8000 and interval_ms % 20 == 0) else 10)
        if prefix:
            self.h = C.c_void_p(L.orc_aec_init(1, freq, interval_ms))
        else:
            self.h = C.c_void_p(L.aec_init(1, freq, interval_ms, None))
        assert self.h

    def process2(self, far, near, delay_ms=0):
        far = np.ascontiguousarray(far, np.int16).copy()
        near = np.ascontiguousarray(near, np.int16).copy()
        out = np.zeros(len(near), np.int16)
        f = self.L.orc_aec_process2 if self.prefix else self.L.aec_process2
        rc = f(self.h, P(far), P(near), P(out), len(near), delay_ms)
        return out, rc

    def set_far(self, far):
        far = np.ascontiguousarray(far, np.int16).copy()
        f = self.L.orc_aec_set_frame_far if self.prefix else self.L.aec_setFrameFar
        return f(self.h, P(far), len(far))

    def process(self, near, delay_ms=0):
        near = np.ascontiguousarray(near, np.int16).copy()
        out = np.zeros(len(near), np.int16)
        f = self.L.orc_aec_process if self.prefix else self.L.aec_process
        rc = f(self.h, P(near), P(out), len(near), delay_ms)
        return out, rc

    def close(self):
        (self.L.orc_aec_release if self.prefix else self.L.aec_release)(self.h)


def aec_run_pairs(L, prefix, far, near, freq, interval_ms=10, delay_ms=0):
    """far/near int16 [T, S, n] -> out [T, S, n] through one checker handle per stream."""
    T, S, n = near.shape
    out = np.empty_like(near)
    for s in range(S):
        h = AecRef(L, freq, interval_ms, prefix)
        for t in range(T):
            d = delay_ms(t) if callable(delay_ms) else delay_ms
            out[t, s], _ = h.process2(far[t, s], near[t, s], d)
        h.close()
    return out


class MixView(C.Structure):
    """orc_mix_view (oracle/oracle.h)"""
    _fields_ = [("ring_bytes", C.c_uint32), ("head_off", C.c_uint32), ("tick", C.c_uint32), ("play_correct", C.c_uint32),
                ("mix_freq", C.c_uint16), ("reduce_mode", C.c_uint8), ("run", C.c_uint8)]


def load_data_cases():
    """(freq, channels, sample, frames, reduce) producers' calls used by the wmix_load_data tests: same format, both resampling
    directions, mono / stereo, the empty 8-bit case"""
    return [(16000, 1, 16, 320, 0), (8000, 1, 16, 160, 3), (44100, 2, 16, 441, 0), (16000, 2, 16, 200, 3), (16000, 1, 8, 100, 0),
            (32000, 1, 16, 640, 1), (11025, 2, 16, 221, 0), (16000, 1, 16, 97, 3)]

