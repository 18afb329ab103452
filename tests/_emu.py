"""Loader for tests/emu/libwmix_emu.so — the CPU lane-loop emulation of the kernel bodies
(test infrastructure; see tests/emu/emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_emu = None


def emu():
    global _emu
    if _emu is None:
        d = os.path.join(ROOT, "tests", "emu")
        so = os.path.join(d, "libwmix_emu.so")
        csrc = os.path.join(ROOT, "wmix_b200", "csrc")
        deps = [os.path.join(d, "emu.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in deps):
            subprocess.check_call(
                ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fwrapv",
                 "-fno-strict-aliasing", "-Wno-unknown-pragmas", "-o", so, os.path.join(d, "emu.cpp"),
                 os.path.join(csrc, "host_tables.cpp"), "-lm"])
        L = C.CDLL(so)
        L.emu_ns_create.restype = C.c_void_p
        L.emu_int_create.restype = C.c_void_p
        L.emu_ns_record.restype = C.POINTER(C.c_float)
        L.emu_nscta_create.restype = C.c_void_p
        L.emu_nscta_record.restype = C.POINTER(C.c_float)
        L.emu_nscta_hist.restype = C.POINTER(C.c_uint16)
        L.emu_ns_hist.restype = C.POINTER(C.c_uint16)
        L.emu_mix_step.restype = C.c_int16
        L.emu_div_by_counter_mismatches.restype = C.c_long
        _emu = L
    return _emu


class EmuChain:
    def __init__(self, freq, ns=True, agc=True, vad=True, gain=5):
        L = emu()
        self.L, self.n = L, freq // 100
        self.ns = C.c_void_p(L.emu_ns_create(freq)) if ns else None
        self.it = C.c_void_p(L.emu_int_create(freq, gain, 3)) if (agc or vad) else None
        self.do_agc, self.do_vad = agc, vad
        self.flags = []

    def frame(self, x):
        x = np.ascontiguousarray(x, dtype=np.int16).copy()
        p = x.ctypes.data_as(C.c_void_p)
        if self.ns:
            self.L.emu_ns_frame(self.ns, p, p)
        if self.do_agc:
            self.L.emu_agc_frame(self.it, p)
        if self.do_vad:
            self.flags.append(self.L.emu_vad_frame(self.it, p))
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
            self.L.emu_ns_destroy(self.ns)
        if self.it:
            self.L.emu_int_destroy(self.it)
