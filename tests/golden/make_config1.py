"""Regenerates the BASELINE config-1 fixtures from the UNMODIFIED reference (oracle/_ref/libwmix_ref.so) and the
reference's own audio fixture R:audio/1x8000.wav.  Run in the build container only:
    python tests/golden/make_config1.py
Writes (raw little-endian int16, mono 8 kHz):
    config1_in_20s.s16       the first 20 s (2000 frames of 80 samples) of audio/1x8000.wav, WAV header stripped
    config1_ns_20s.s16       ns_init(1, 8000, NULL) + ns_process over those frames (wmix's own single-stream NS)
    config1_chain_20s.s16    the same frames through NS -> AGC(5) -> VAD(10 ms) in place
The fixtures travel with the repo; the reference does not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests._oracle import RefChain, ref  # noqa: E402

FRAMES = 2000


def main():
    R = ref()
    assert R is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
    here = os.path.dirname(os.path.abspath(__file__))
    pcm = np.fromfile("/root/reference/audio/1x8000.wav", dtype=np.int16, offset=44)[: FRAMES * 80]
    pcm.tofile(os.path.join(here, "config1_in_20s.s16"))
    a = RefChain(R, 8000, ns=True, agc=False, vad=False)
    a.run(pcm).tofile(os.path.join(here, "config1_ns_20s.s16"))
    a.close()
    b = RefChain(R, 8000)
    b.run(pcm).tofile(os.path.join(here, "config1_chain_20s.s16"))
    b.close()
    print("wrote 3 x %d samples" % len(pcm))


if __name__ == "__main__":
    main()
