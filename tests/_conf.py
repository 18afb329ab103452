"""Conference-bus checker for the tests: the reference's G.711 tables (through the oracle) and a numpy
statement of "exact int32 bus, N-minus-one read-out".  Also a numpy stand-in for the CUDA backend of
wmix_b200.conference so the multi-rank host logic (placement, exchange wiring) can run on CPU over gloo.
Test infrastructure only."""
import numpy as np

from tests._oracle import oracle

_tabs = {}


def g711_tables(law):
    """(decode[256] int16, encode[65536] uint8 indexed by pcm & 0xFFFF) from the oracle's per-sample codecs"""
    if law not in _tabs:
        L = oracle()
        dec_f = L.orc_alaw2linear if law == 0 else L.orc_ulaw2linear
        enc_f = L.orc_linear2alaw if law == 0 else L.orc_linear2ulaw
        dec = np.array([dec_f(c) for c in range(256)], np.int16)
        enc = np.zeros(65536, np.uint8)
        for v in range(-32768, 32768):
            enc[v & 0xFFFF] = enc_f(v)
        _tabs[law] = (dec, enc)
    return _tabs[law]


def decode(law, x):
    return x.astype(np.int32) if law < 0 else g711_tables(law)[0][x].astype(np.int32)


def encode(law, pcm16):
    return pcm16.astype(np.int16) if law < 0 else g711_tables(law)[1][pcm16.astype(np.int32) & 0xFFFF]


def conference_oracle(law, legs, conf_start):
    """legs [P, frame] (codes or int16), conf_start [n_conf+1] -> (bus int32 [n_conf, frame], out like legs)"""
    pcm = decode(law, legs)
    n_conf = len(conf_start) - 1
    bus = np.zeros((n_conf, legs.shape[1]), np.int32)
    out = np.empty_like(legs)
    for c in range(n_conf):
        a, b = conf_start[c], conf_start[c + 1]
        bus[c] = pcm[a:b].sum(axis=0)
        out[a:b] = encode(law, np.clip(bus[c][None, :] - pcm[a:b], -32768, 32767))
    return bus, out


class NumpyBackend:
    """Stand-in for wmix_b200.conference.CudaBackend on CPU tensors (torch CPU or numpy)."""

    def __init__(self):
        self.conf_start = None

    @staticmethod
    def _np(x):
        return x if isinstance(x, np.ndarray) else x.numpy()

    def set_conferences(self, conf_start):
        self.conf_start = np.asarray(conf_start)

    def bus_sum(self, law, d_in, d_bus, stream):
        pcm = decode(law, self._np(d_in))
        bus = self._np(d_bus)
        for c in range(len(self.conf_start) - 1):
            bus[c] = pcm[self.conf_start[c]:self.conf_start[c + 1]].sum(axis=0)

    def nminus1(self, law, d_bus, d_in, d_out, stream):
        legs, bus, out = self._np(d_in), self._np(d_bus), self._np(d_out)
        pcm = decode(law, legs)
        for c in range(len(self.conf_start) - 1):
            a, b = self.conf_start[c], self.conf_start[c + 1]
            out[a:b] = encode(law, np.clip(bus[c][None, :] - pcm[a:b], -32768, 32767))

    def close(self):
        pass
