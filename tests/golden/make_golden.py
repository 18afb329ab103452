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
    json.dump(g, open(os.path.join(ROOT, "tests", "golden", "hashes.json"), "w"), indent=1)
    print("wrote hashes.json")


if __name__ == "__main__":
    main()
