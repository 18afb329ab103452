"""Regenerates tests/golden/hashes.json from the UNMODIFIED reference (oracle/_ref/libwmix_ref.so,
built by oracle/build_ref.sh from /root/reference).  Run in the build container only:
    python tests/golden/make_golden.py
The fixtures travel with the repo; the reference does not."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests._oracle import P, RefChain, fnv1a64, ref  # noqa: E402
from wmix_b200.synth import make_frames  # noqa: E402


def main():
    R = ref()
    assert R is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
    g = {}
    x = np.arange(-32768, 32768, dtype=np.int16)
    o = np.zeros(65536, np.uint8)
    R.PCM2G711a(P(x), P(o), 131072, 0)
    g["g711_alaw_enc"] = fnv1a64(o.tobytes())
    R.PCM2G711u(P(x), P(o), 131072, 0)
    g["g711_ulaw_enc"] = fnv1a64(o.tobytes())
    c = np.arange(256, dtype=np.uint8)
    d = np.zeros(256, np.int16)
    R.G711a2PCM(P(c), P(d), 256, 0)
    g["g711_alaw_dec"] = fnv1a64(d.tobytes())
    R.G711u2PCM(P(c), P(d), 256, 0)
    g["g711_ulaw_dec"] = fnv1a64(d.tobytes())

    wav = "/root/reference/audio/1x8000.wav"
    pcm = np.fromfile(wav, dtype=np.int16, offset=44)
    pcm = pcm[: len(pcm) // 80 * 80]
    ch = RefChain(R, 8000)
    g["config1_ns_agc_vad"] = fnv1a64(ch.run(pcm).tobytes())
    ch.close()

    g["streams"] = {}
    for freq in (8000, 16000):
        for stage, S, T in (("vad", 4, 300), ("agc", 4, 300), ("ns", 3, 650), ("chain", 3, 650)):
            seed = 21
            xs = make_frames(S, freq, 0, T, seed=seed)
            kw = dict(ns=stage in ("ns", "chain"), agc=stage in ("agc", "chain"), vad=stage in ("vad", "chain"))
            outs = []
            for s in range(S):
                c_ = RefChain(R, freq, **kw)
                outs.append(c_.run(np.ascontiguousarray(xs[:, s, :]).reshape(-1)))
                c_.close()
            y = np.stack(outs)
            g["streams"]["%s_%d" % (stage, freq)] = dict(freq=freq, stage=stage, n_streams=S, n_ticks=T, seed=seed,
                                                         hash=fnv1a64(y.tobytes()),
                                                         head=y[:, :8].tolist(), tail=y[:, -8:].tolist())
    # AEC: aec_process2 on near/far pairs (config 4's signal model)
    from tests._oracle import aec_run_pairs
    from wmix_b200.synth import make_aec_pairs

    g["aec"] = {}
    for freq, T, delay in ((8000, 900, 0), (16000, 500, 0), (8000, 500, 80)):
        S, seed = 4, 29
        far, near = make_aec_pairs(S, freq, 0, T, seed=seed)
        y = aec_run_pairs(R, "", far, near, freq, 10, delay)
        y = np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(S, -1)
        g["aec"]["aec_%d_d%d" % (freq, delay)] = dict(freq=freq, n_streams=S, n_ticks=T, seed=seed, delay_ms=delay,
                                                     hash=fnv1a64(y.tobytes()), tail=y[:, -8:].tolist())
    # handle-layer quirks at 32 kHz and the stereo "right channel as high band" NS (R:src/webrtc.c:624-636, :727)
    import ctypes as C

    g["handles"] = {}
    S, T, seed = 2, 300, 33
    xs = make_frames(2 * S, 16000, 0, 2 * T, seed=seed)
    for stage in ("ns", "agc", "vad", "chain"):
        kw = dict(ns=stage in ("ns", "chain"), agc=stage in ("agc", "chain"), vad=stage in ("vad", "chain"))
        outs = []
        for s in range(S):
            c_ = RefChain(R, 32000, **kw)
            outs.append(c_.run(np.ascontiguousarray(xs[:, s, :]).reshape(-1)))
            c_.close()
        y = np.stack(outs)
        g["handles"]["%s_32000" % stage] = dict(freq=32000, stage=stage, n_streams=S, n_ticks=T, seed=seed, hash=fnv1a64(y.tobytes()),
                                                tail=y[:, -8:].tolist())
    for freq in (8000, 16000):
        n = freq // 100
        xf = make_frames(4, freq, 0, T, seed=seed)
        outs = []
        for a, b in ((0, 3), (1, 2)):
            h = C.c_void_p(R.ns_init(2, freq, None))
            st = np.empty((T, 2 * n), np.int16)
            st[:, 0::2], st[:, 1::2] = xf[:, a], xf[:, b]
            out = np.zeros_like(st)
            for t in range(T):
                R.ns_process(h, P(st[t].copy()), P(out[t]), n)
            R.ns_release(h)
            outs.append(out.reshape(-1))
        y = np.stack(outs)
        g["handles"]["ns_stereo_%d" % freq] = dict(freq=freq, pairs=[[0, 3], [1, 2]], n_ticks=T, seed=seed, hash=fnv1a64(y.tobytes()),
                                                   tail=y[:, -8:].tolist())

    # resampling branches of the real wmix_load_data into the 16 kHz mono bus of the oracle build (R:src/wmix.c:1704-1939)
    class WPoint(C.Union):
        _fields_ = [("U8", C.c_void_p)]

    R.wmix_load_data.restype = WPoint
    R.wmix_load_data.argtypes = [C.c_void_p, WPoint, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, WPoint, C.c_uint8, C.POINTER(C.c_uint32)]
    wm = (C.c_uint8 * R.oracle_ref_sizeof_wmix())()
    ring_bytes = R.oracle_ref_wmix_buff_size()
    n = ring_bytes // 2
    g["mix_resample"] = dict(mix_freq=R.oracle_ref_wmix_freq(), ring_samples=n, seed=35, cases=[])
    rng = np.random.default_rng(35)
    for freq, chn, rdce in ((8000, 1, 1), (8000, 2, 3), (11025, 1, 16), (44100, 2, 1), (48000, 1, 3), (16000, 2, 1)):
        ring = rng.integers(-32768, 32768, n).astype(np.int16)
        src = rng.integers(-32768, 32768, 401 * chn + 2).astype(np.int16)
        seed_ring, seed_src = fnv1a64(ring.tobytes()), fnv1a64(src.tobytes())
        head_off = (n - 123) * 2
        R.oracle_ref_wmix_seat(wm, P(ring), ring_bytes, rdce, 0, 0)
        tick = C.c_uint32(0)
        out = R.wmix_load_data(wm, WPoint(src.ctypes.data), 401 * chn * 2, freq, chn, 16, WPoint(ring.ctypes.data + head_off),
                               0 if rdce != 1 else 1, C.byref(tick))
        g["mix_resample"]["cases"].append(dict(freq=freq, chn=chn, rdce=rdce, frames=401, head=n - 123, ring_in=seed_ring, src_in=seed_src,
                                               ring_out=fnv1a64(ring.tobytes()), new_head=(out.U8 - ring.ctypes.data) // 2,
                                               written=tick.value // 2))

    # playPkgBuff_get's slot for every write index (R:src/wmix.c:496-509) with the oracle build's ring (AEC_INTERVALMS = 400)
    pkg, num = R.oracle_ref_wmix_pkg_size(), R.oracle_ref_wmix_aec_fifo_pkgs()
    R.playPkgBuff_get.restype = C.c_void_p
    R.playPkgBuff_get.argtypes = [C.c_void_p, C.c_int]
    for _ in range(num):
        R.playPkgBuff_add(P(np.zeros(pkg, np.uint8)))
    seq = []
    for t in range(3 * num):
        R.playPkgBuff_add(P(np.full(pkg, (t % 250) + 1, np.uint8)))
        b = np.zeros(pkg, np.uint8)
        R.playPkgBuff_get(b.ctypes.data, 400)
        seq.append(int(b[0]))                      # 1 + index of the add this package came from (0 = never written)
    for _ in range(num):
        R.playPkgBuff_add(P(np.zeros(pkg, np.uint8)))
    g["play_fifo"] = dict(n_pkg=num, delay_pkgs=400 // R.oracle_ref_wmix_interval_ms(), got_tag=seq)
    json.dump(g, open(os.path.join(ROOT, "tests", "golden", "hashes.json"), "w"), indent=1)
    print("wrote hashes.json")


if __name__ == "__main__":
    main()
