"""Seeded synthetic speech-plus-noise streams, vectorised over streams.

Host-side input generator used by bench.py and the tests.  Stream `s` is reproducible from
(seed, s) alone: a voiced source A*(sin p + 0.5 sin 2p + 0.3 sin 3p) with slow vibrato, gated
by talk-spurts, plus always-on uniform white noise; every 64th stream is a corner-case cohort
(all-zero, full-scale square, DC offset, silence-then-speech step)."""
import numpy as np


def _lcg(x):
    return (x * np.uint64(1664525) + np.uint64(1013904223)) & np.uint64(0xFFFFFFFF)


def stream_params(n_streams, seed=0):
    s = np.arange(n_streams, dtype=np.uint64)
    x = (np.uint64(0x9E3779B9) * (s + np.uint64(1)) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    out = []
    for _ in range(8):
        x = _lcg(x)
        out.append((x >> np.uint64(8)).astype(np.float64) / float(1 << 24))
    u = np.stack(out, axis=0)
    return dict(
        f0=90.0 + 160.0 * u[0],
        amp=1500.0 + 10500.0 * u[1],
        period=0.6 + 1.4 * u[2],
        phase=u[3],
        noise_db=-45.0 + 20.0 * u[4],
        vib=0.5 + 3.0 * u[5],
        duty=0.35 + 0.4 * u[6],
        nseed=(x & np.uint64(0xFFFFFFFF)).astype(np.uint64),
    )


def make_frames(n_streams, freq, tick0, n_ticks, seed=0, cohorts=True):
    """int16 array [n_ticks, n_streams, freq//100] for ticks tick0 .. tick0+n_ticks-1."""
    L = freq // 100
    p = stream_params(n_streams, seed)
    t = (tick0 * L + np.arange(n_ticks * L, dtype=np.float64)) / freq          # [T]
    tt = t[None, :]
    f0 = p["f0"][:, None] * (1.0 + 0.02 * np.sin(2 * np.pi * p["vib"][:, None] * tt))
    # phase by closed form of the vibrato integral keeps ticks independent of each other
    ph = 2 * np.pi * (p["f0"][:, None] * tt - 0.02 * p["f0"][:, None] / (2 * np.pi * p["vib"][:, None])
                      * (np.cos(2 * np.pi * p["vib"][:, None] * tt) - 1.0))
    del f0
    voiced = np.sin(ph) + 0.5 * np.sin(2 * ph) + 0.3 * np.sin(3 * ph)
    gate = (((tt / p["period"][:, None]) + p["phase"][:, None]) % 1.0) < p["duty"][:, None]
    sig = p["amp"][:, None] * voiced * gate
    # white noise from a counter-based hash so any tick can be generated on its own
    idx = (tick0 * L + np.arange(n_ticks * L, dtype=np.uint64))[None, :]
    h = (idx * np.uint64(2654435761) + p["nseed"][:, None] * np.uint64(40503)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(3266489917)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    un = h.astype(np.float64) / float(1 << 32) * 2.0 - 1.0
    namp = 32768.0 * 10.0 ** (p["noise_db"][:, None] / 20.0) * np.sqrt(3.0)
    x = sig + namp * un
    if cohorts and n_streams >= 4:
        sid = np.arange(n_streams)
        x[sid % 64 == 1] = 0.0
        sq = np.where((np.arange(n_ticks * L) + tick0 * L) // 40 % 2 == 0, 32767.0, -32767.0)
        x[sid % 64 == 2] = sq[None, :]
        x[sid % 64 == 3] += 6000.0
        late = (sid % 64 == 0)
        early = (t < 2.5)
        x[np.ix_(late, early)] = 0.0
    x = np.clip(np.rint(x), -32768, 32767).astype(np.int16)
    return np.ascontiguousarray(x.reshape(n_streams, n_ticks, L).transpose(1, 0, 2))


def make_aec_pairs(n_streams, freq, tick0, n_ticks, seed=0, delay_ms=5, echo_gain=0.5, cohorts=True):
    """(far, near) int16 arrays [n_ticks, n_streams, freq//100] for BASELINE config 4:
    near = far delayed by `delay_ms` and scaled by `echo_gain` + an independent local talker + noise.
    Cohorts (s mod 64): 1 = far all-zero (zero far-end guard), 2 = far full-scale square (saturating echo)."""
    L = freq // 100
    d = freq * delay_ms // 1000
    lead = (d + L - 1) // L
    t0 = max(0, tick0 - lead)
    far_all = make_frames(n_streams, freq, t0, n_ticks + (tick0 - t0), seed=seed, cohorts=cohorts)
    far_flat = far_all.transpose(1, 0, 2).reshape(n_streams, -1).astype(np.float64)
    pad = lead * L - (tick0 - t0) * L                       # zeros before the beginning of time
    far_flat = np.concatenate([np.zeros((n_streams, pad)), far_flat], axis=1)
    start = lead * L
    echo = echo_gain * far_flat[:, start - d:start - d + n_ticks * L]
    local = make_frames(n_streams, freq, tick0, n_ticks, seed=seed + 7919, cohorts=False).transpose(1, 0, 2)
    local = local.reshape(n_streams, -1).astype(np.float64) * 0.6
    near = np.clip(np.rint(echo + local), -32768, 32767).astype(np.int16)
    far = far_flat[:, start:start + n_ticks * L].astype(np.int16)
    shp = (n_streams, n_ticks, L)
    return (np.ascontiguousarray(far.reshape(shp).transpose(1, 0, 2)),
            np.ascontiguousarray(near.reshape(shp).transpose(1, 0, 2)))
