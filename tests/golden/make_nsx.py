"""Regenerates tests/golden/nsx.json from the UNMODIFIED reference: the fixed-point suppressor WebRtcNsx_* of
oracle/_ref/libwmix_ref.so, the handle layer built with the reference's own switch (oracle/_ref/libwmix_ref_nsx.so =
R:src/webrtc.c compiled with -DMAKE_WEBRTC_NSX, oracle/build_ref.sh) and the literal tables of T:.../ns/nsx_core.c.
Run in the build container only:   python tests/golden/make_nsx.py"""
import json
import os
import re
import sys
import tarfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests._oracle import NsxCore, fnv1a64, nsx_handle_run, nsx_quiet_streams, ref, ref_nsx  # noqa: E402
from wmix_b200.synth import make_frames  # noqa: E402

TAR = "/root/reference/pkg/webrtc_cut.tar.gz"
NS_DIR = "webrtc_cut/webrtc/modules/audio_processing/ns/"
SPL_DIR = "webrtc_cut/webrtc/common_audio/signal_processing/"
# the order of oracle/orc_nsx.c's table dump (orc_nsx_tables)
TABLES = [("nsx_core.c", "WebRtcNsx_kLogTable"), ("nsx_core.c", "WebRtcNsx_kCounterDiv"), ("nsx_core.c", "WebRtcNsx_kLogTableFrac"),
          ("nsx_core.c", "kBlocks80w128x"), ("nsx_core.c", "kBlocks160w256x"), ("nsx_core.c", "kFactor1Table"),
          ("nsx_core.c", "kFactor2Aggressiveness1"), ("nsx_core.c", "kFactor2Aggressiveness2"), ("nsx_core.c", "kFactor2Aggressiveness3"),
          ("nsx_core.c", "kSumLogIndex"), ("nsx_core.c", "kSumSquareLogIndex"), ("nsx_core.c", "kLogIndex"),
          ("nsx_core.c", "kDeterminantEstMatrix"), ("complex_fft_tables.h", "kSinTable1024"), ("nsx_core_c.c", "kIndicatorTable")]


def reference_tables():
    """the literal int16 tables of the reference, concatenated in TABLES order"""
    tf = tarfile.open(TAR)
    src = {}
    for f in ("nsx_core.c", "nsx_core_c.c"):
        src[f] = tf.extractfile(NS_DIR + f).read().decode()
    src["complex_fft_tables.h"] = tf.extractfile(SPL_DIR + "complex_fft_tables.h").read().decode()
    out = []
    for f, name in TABLES:
        m = re.search(name + r"\[\d*\]\s*=\s*\{([^}]*)\}", src[f])
        out.append(np.array([int(x) for x in re.findall(r"-?\d+", m.group(1))], np.int16))
    return out


def streams_through_core(L, freq, pcm, policy=2, hb=None):
    T, S, n = pcm.shape
    out = np.zeros_like(pcm)
    out_hb = np.zeros_like(pcm) if hb is not None else None
    for s in range(S):
        c = NsxCore(L, freq, policy)
        for t in range(T):
            if hb is None:
                out[t, s] = c.frame(pcm[t, s])
            else:
                out[t, s], out_hb[t, s] = c.frame(pcm[t, s], hb[t, s])
        c.close()
    return out if hb is None else (out, out_hb)


def describe(y):
    y = np.ascontiguousarray(y)
    flat = y.reshape(-1)
    return dict(hash=fnv1a64(y.tobytes()), head=flat[:8].tolist(), tail=flat[-8:].tolist(), shape=list(y.shape))


def main():
    R, RX = ref(), ref_nsx()
    assert R is not None and RX is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
    g = {"tables": {}, "core": {}, "handle": {}}
    for (f, name), t in zip(TABLES, reference_tables()):
        g["tables"][name] = dict(n=int(len(t)), hash=fnv1a64(t.tobytes()))
    for freq in (8000, 16000):
        pcm = make_frames(4, freq, 0, 1100, seed=21)
        for policy in (0, 1, 2, 3):
            g["core"]["synth_%d_p%d" % (freq, policy)] = dict(freq=freq, policy=policy, n_streams=4, n_ticks=1100, seed=21,
                                                              **describe(streams_through_core(R, freq, pcm, policy)))
        q = nsx_quiet_streams(freq)
        g["core"]["quiet_%d" % freq] = dict(freq=freq, policy=2, **describe(streams_through_core(R, freq, q)))
        hb = make_frames(4, freq, 0, 1100, seed=22)
        lo, hi = streams_through_core(R, freq, pcm, 2, hb)
        g["core"]["stereo_%d" % freq] = dict(freq=freq, policy=2, seed=21, seed_hb=22, lo=describe(lo), hi=describe(hi))
    # the wrapper with the switch thrown: mono and stereo handles at 8 / 16 / 32 kHz (32 kHz: 320-sample packets)
    for freq in (8000, 16000, 32000):
        for chn in (1, 2):
            x = make_frames(chn, freq, 0, 300, seed=23)                       # [T, chn, n] -> interleave
            pcm = np.ascontiguousarray(x.transpose(0, 2, 1).reshape(300, -1))
            g["handle"]["%d_%d" % (chn, freq)] = dict(chn=chn, freq=freq, n_ticks=300, seed=23,
                                                      **describe(nsx_handle_run(RX, "", chn, freq, pcm)))
    wav = np.fromfile(os.path.join(ROOT, "tests", "golden", "config1_in_20s.s16"), np.int16)
    g["config1_nsx"] = describe(nsx_handle_run(RX, "", 1, 8000, wav.reshape(-1, 80)))
    json.dump(g, open(os.path.join(ROOT, "tests", "golden", "nsx.json"), "w"), indent=1)
    print("wrote tests/golden/nsx.json")


if __name__ == "__main__":
    main()
