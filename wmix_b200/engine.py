"""Python face of the batched engine (include/wmixb.h): thin, pointer-passing wrappers.

PyTorch is used here only for device memory and streams (tensors are handed to the C-ABI as raw
pointers); numpy arrays go through the host-buffer entry point."""
import ctypes as C

import numpy as np

from ._lib import AEC, AGC, NS, VAD, Config, check, lib  # noqa: F401


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


def _stream_ptr(stream):
    if stream is None:
        import torch

        return torch.cuda.current_stream().cuda_stream
    return getattr(stream, "cuda_stream", stream)


class Engine:
    """N independent mono streams of wmix's record chain NS -> AGC -> VAD on one GPU."""

    def __init__(self, n_streams, freq=16000, stages=NS | AGC | VAD, ns_policy=2, agc_gain_db=5, vad_mode=3, device=0,
                 aec_far_depth=0, ns_high_band=0, ns_core=0):
        self.L = lib()
        cfg = Config(n_streams=n_streams, freq=freq, stages=stages, ns_policy=ns_policy, agc_gain_db=agc_gain_db,
                     vad_mode=vad_mode, device=device, aec_far_depth=aec_far_depth, ns_high_band=ns_high_band,
                     ns_core=ns_core)
        h = C.c_void_p()
        check(self.L.wmixb_create(C.byref(cfg), C.byref(h)), "wmixb_create")
        self.h, self.n, self.freq, self.frame, self.stages, self.device = h, n_streams, freq, freq // 100, stages, device

    def close(self):
        if self.h:
            self.L.wmixb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- ticks ---
    def tick_device(self, d_in, d_out, d_vad=None, stages=0, stream=None):
        check(self.L.wmixb_tick_device(self.h, _ptr(d_in), _ptr(d_out), _ptr(d_vad), stages, _stream_ptr(stream)),
              "wmixb_tick_device")

    def tick_host(self, h_in, h_out, h_vad=None, stages=0):
        check(self.L.wmixb_tick_host(self.h, _ptr(h_in), _ptr(h_out), _ptr(h_vad), stages), "wmixb_tick_host")

    def tick_host_submit(self, h_in, h_out, h_vad=None, h_bus=None, stages=0):
        """queue one tick (see wmixb_tick_host_submit); keep the buffers alive until tick_host_wait() returns for it"""
        check(self.L.wmixb_tick_host_submit(self.h, _ptr(h_in), _ptr(h_out), _ptr(h_vad), _ptr(h_bus), stages),
              "wmixb_tick_host_submit")

    def tick_host_wait(self):
        check(self.L.wmixb_tick_host_wait(self.h), "wmixb_tick_host_wait")

    def tick_host_bus(self, h_in, h_out, h_vad, h_bus, stages=0):
        check(self.L.wmixb_tick_host_bus(self.h, _ptr(h_in), _ptr(h_out), _ptr(h_vad), _ptr(h_bus), stages),
              "wmixb_tick_host_bus")

    def vad20_device(self, d_pcm, d_vad=None, stream=None):
        check(self.L.wmixb_vad20_device(self.h, _ptr(d_pcm), _ptr(d_vad), _stream_ptr(stream)), "wmixb_vad20_device")

    def vad20_host(self, h_pcm, h_vad=None):
        check(self.L.wmixb_vad20_host(self.h, _ptr(h_pcm), _ptr(h_vad)), "wmixb_vad20_host")

    def offline_device(self, d_in, d_out, n_frames, d_vad=None, stages=0, stream=None):
        check(self.L.wmixb_offline_device(self.h, _ptr(d_in), _ptr(d_out), _ptr(d_vad), n_frames, stages,
                                          _stream_ptr(stream)), "wmixb_offline_device")

    # --- echo canceller ---
    def aec_device(self, d_far, d_near, d_out, samples=None, delay_ms=0, stream=None):
        check(self.L.wmixb_aec_device(self.h, _ptr(d_far), _ptr(d_near), _ptr(d_out), samples or self.frame, delay_ms,
                                      _stream_ptr(stream)), "wmixb_aec_device")

    def aec_host(self, h_far, h_near, h_out, samples=None, delay_ms=0):
        check(self.L.wmixb_aec_host(self.h, _ptr(h_far), _ptr(h_near), _ptr(h_out), samples or self.frame, delay_ms),
              "wmixb_aec_host")

    def tick_chain_device(self, d_far, d_in, d_out, d_vad=None, stages=0, delay_ms=0, stream=None):
        check(self.L.wmixb_tick_chain_device(self.h, _ptr(d_far), _ptr(d_in), _ptr(d_out), _ptr(d_vad), stages, delay_ms,
                                             _stream_ptr(stream)), "wmixb_tick_chain_device")

    def aec_status(self):
        """(OR of the sticky per-stream AEC flags, number of flagged streams)"""
        flags, n = C.c_int(0), C.c_int(0)
        check(self.L.wmixb_aec_status(self.h, C.byref(flags), C.byref(n)), "wmixb_aec_status")
        return flags.value, n.value

    # --- conference bus ---
    def set_conferences(self, conf_start):
        a = np.ascontiguousarray(conf_start, dtype=np.int32)
        check(self.L.wmixb_set_conferences(self.h, a.ctypes.data, len(a) - 1), "wmixb_set_conferences")
        self.n_conf = len(a) - 1

    def bus_sum(self, d_pcm, d_bus, stream=None):
        check(self.L.wmixb_bus_sum_device(self.h, _ptr(d_pcm), _ptr(d_bus), _stream_ptr(stream)), "wmixb_bus_sum_device")

    def bus_nminus1(self, d_bus, d_pcm, d_out, stream=None):
        check(self.L.wmixb_bus_nminus1_device(self.h, _ptr(d_bus), _ptr(d_pcm), _ptr(d_out), _stream_ptr(stream)),
              "wmixb_bus_nminus1_device")

    def g711_bus_sum(self, law, d_codes, d_bus, stream=None):
        check(self.L.wmixb_g711_bus_sum_device(self.h, law, _ptr(d_codes), _ptr(d_bus), _stream_ptr(stream)),
              "wmixb_g711_bus_sum_device")

    def g711_nminus1(self, law, d_bus, d_codes, d_out, stream=None):
        check(self.L.wmixb_g711_nminus1_device(self.h, law, _ptr(d_bus), _ptr(d_codes), _ptr(d_out), _stream_ptr(stream)),
              "wmixb_g711_nminus1_device")

    # --- misc ---
    def reset(self, first=0, count=None):
        check(self.L.wmixb_reset(self.h, first, self.n - first if count is None else count), "wmixb_reset")

    def set_agc_gain(self, db):
        check(self.L.wmixb_set_agc_gain(self.h, db), "wmixb_set_agc_gain")

    def sync(self):
        check(self.L.wmixb_sync(self.h), "wmixb_sync")

    def set_tuning(self, key, value):
        """experiment / test knobs of include/wmixb.h (none changes results)"""
        check(self.L.wmixb_set_tuning(self.h, key.encode(), int(value)), "wmixb_set_tuning(%s)" % key)

    def state_bytes_per_stream(self):
        return self.L.wmixb_state_bytes_per_stream(self.h)

    def get_state(self, s):
        buf = np.zeros(self.L.wmixb_stream_state_bytes(self.h), np.uint8)
        check(self.L.wmixb_get_stream_state(self.h, s, buf.ctypes.data), "wmixb_get_stream_state")
        return buf

    def set_state(self, s, buf):
        buf = np.ascontiguousarray(buf, np.uint8)
        check(self.L.wmixb_set_stream_state(self.h, s, buf.ctypes.data), "wmixb_set_stream_state")


def g711_encode(law, d_pcm, d_codes, n, stream=None):
    check(lib().wmixb_g711_encode_device(law, _ptr(d_pcm), _ptr(d_codes), n, _stream_ptr(stream)), "wmixb_g711_encode_device")


def g711_decode(law, d_codes, d_pcm, n, stream=None):
    check(lib().wmixb_g711_decode_device(law, _ptr(d_codes), _ptr(d_pcm), n, _stream_ptr(stream)), "wmixb_g711_decode_device")


def mix_load(d_ring, ring_len, pos, d_src, n, rdce, stream=None):
    new_pos = C.c_uint32(0)
    check(lib().wmixb_mix_load_device(_ptr(d_ring), ring_len, pos, _ptr(d_src), n, rdce, C.byref(new_pos),
                                      _stream_ptr(stream)), "wmixb_mix_load_device")
    return new_pos.value


class HostBuffer:
    """Pinned host memory from wmixb_host_alloc (NUMA-local to `device`), viewed as a numpy array."""

    def __init__(self, shape, dtype, device=0, write_combined=False):
        self.L = lib()
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = self.L.wmixb_host_alloc(self.nbytes, device, 1 if write_combined else 0)
        if not self.ptr:
            raise RuntimeError("wmixb_host_alloc(%d bytes) failed: %s" % (self.nbytes, self.L.wmixb_last_error().decode()))
        buf = (C.c_ubyte * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.wmixb_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_copy_ceiling(device, h_src, h_dst, h2d_bytes, d2h_bytes, reps=20):
    """ms per repetition of bare cudaMemcpyAsync traffic (wmixb_host_copy_ceiling)"""
    ms = C.c_double(0.0)
    check(lib().wmixb_host_copy_ceiling(device, _ptr(h_src), _ptr(h_dst), h2d_bytes, d2h_bytes, reps, C.byref(ms)),
          "wmixb_host_copy_ceiling")
    return ms.value


def kernel_launches():
    return lib().wmixb_kernel_launches()
