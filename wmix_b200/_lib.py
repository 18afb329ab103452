"""ctypes binding of libwmix_b200.so (the C-ABI in include/wmixb.h, include/webrtc.h,
include/g711codec.h).  There is no Python or CPU implementation behind this module: if the
shared library is missing or cannot be loaded, importing a symbol raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WMIX_B200_LIB: load another build of the same library (kernel experiments); default = the in-tree build
LIB_PATH = os.environ.get("WMIX_B200_LIB") or os.path.join(_HERE, "libwmix_b200.so")

NS, AGC, VAD, AEC = 1, 2, 4, 8


class Config(C.Structure):
    _fields_ = [("n_streams", C.c_int), ("freq", C.c_int), ("stages", C.c_int), ("ns_policy", C.c_int),
                ("agc_gain_db", C.c_int), ("vad_mode", C.c_int), ("device", C.c_int), ("aec_far_depth", C.c_int),
                ("ns_high_band", C.c_int), ("ns_core", C.c_int), ("reserved", C.c_int * 6)]


class MixView(C.Structure):
    """wmixb_mix_view (include/wmixb.h)"""
    _fields_ = [("ring_start", C.c_void_p), ("ring_bytes", C.c_uint32), ("head_off", C.c_uint32), ("tick", C.c_uint32),
                ("play_correct", C.c_uint32), ("mix_freq", C.c_uint16), ("reduce_mode", C.c_uint8), ("run", C.c_uint8),
                ("device", C.c_int)]


class WMixStructPrefix(C.Structure):
    """leading fields of the daemon's WMix_Struct as include/wmix.h restates them (R:src/wmixConf.h:176-207)"""
    _fields_ = [("objAo", C.c_void_p), ("objAi", C.c_void_p), ("buff", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p),
                ("head", C.c_void_p), ("tail", C.c_void_p), ("run", C.c_bool), ("loopWord", C.c_uint8),
                ("loopWordRecord", C.c_uint8), ("loopWordFifo", C.c_uint8), ("loopWordRtp", C.c_uint8), ("tick", C.c_uint32),
                ("thread_sys", C.c_uint32), ("thread_record", C.c_uint32), ("thread_play", C.c_uint32), ("playRun", C.c_bool),
                ("recordRun", C.c_bool), ("shmemRun", C.c_int), ("msg_key", C.c_int), ("msg_fd", C.c_int),
                ("reduceMode", C.c_uint8)]


class PeerOpts(C.Structure):
    """wmixb_peer_opts (include/wmixb.h)"""
    _fields_ = [("tile", C.c_int), ("reduce_scatter", C.c_int), ("timeout_ms", C.c_int), ("ranks_per_device", C.c_int),
                ("reserved", C.c_int * 4)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "wmix_b200: %s not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    sig = {
        "wmixb_create": (i, [C.POINTER(Config), C.POINTER(vp)]),
        "wmixb_destroy": (None, [vp]),
        "wmixb_reset": (i, [vp, i, i]),
        "wmixb_set_agc_gain": (i, [vp, i]),
        "wmixb_tick_device": (i, [vp, vp, vp, vp, i, vp]),
        "wmixb_tick_host": (i, [vp, vp, vp, vp, i]),
        "wmixb_tick_host_bus": (i, [vp, vp, vp, vp, vp, i]),
        "wmixb_tick_host_submit": (i, [vp, vp, vp, vp, vp, i]),
        "wmixb_tick_host_wait": (i, [vp]),
        "wmixb_vad20_device": (i, [vp, vp, vp, vp]),
        "wmixb_vad20_host": (i, [vp, vp, vp]),
        "wmixb_record_create": (i, [vp, i, C.POINTER(vp)]),
        "wmixb_record_destroy": (None, [vp]),
        "wmixb_record_far_slot": (i, [vp]),
        "wmixb_record_tick_device": (i, [vp, vp, vp, vp, vp, vp, i, vp]),
        "wmixb_record_tick_host": (i, [vp, vp, vp, vp, vp, i]),
        "wmixb_ns2_device": (i, [vp, vp, vp, vp, vp, vp]),
        "wmixb_ns2_host": (i, [vp, vp, vp, vp, vp]),
        "wmixb_vad32_device": (i, [vp, vp, vp, vp]),
        "wmixb_vad32_host": (i, [vp, vp, vp]),
        "wmixb_offline_device": (i, [vp, vp, vp, vp, i, i, vp]),
        "wmixb_aec_device": (i, [vp, vp, vp, vp, i, i, vp]),
        "wmixb_aec_host": (i, [vp, vp, vp, vp, i, i]),
        "wmixb_tick_chain_device": (i, [vp, vp, vp, vp, vp, i, i, vp]),
        "wmixb_aec_status": (i, [vp, C.POINTER(i), C.POINTER(i)]),
        "wmixb_set_conferences": (i, [vp, vp, i]),
        "wmixb_bus_sum_device": (i, [vp, vp, vp, vp]),
        "wmixb_bus_nminus1_device": (i, [vp, vp, vp, vp, vp]),
        "wmixb_peer_bus_create": (i, [vp, i, i, C.POINTER(vp)]),
        "wmixb_peer_bus_create_ex": (i, [vp, i, i, C.POINTER(PeerOpts), C.POINTER(vp)]),
        "wmixb_peer_bus_destroy": (None, [vp]),
        "wmixb_peer_bus_handle": (i, [vp, vp]),
        "wmixb_peer_bus_connect": (i, [vp, vp]),
        "wmixb_peer_bus_connect_local": (i, [vp, C.POINTER(vp)]),
        "wmixb_peer_bus_tick_device": (i, [vp, i, vp, vp, vp, vp]),
        "wmixb_peer_bus_status": (i, [vp, C.POINTER(i)]),
        "wmixb_g711_encode_device": (i, [i, vp, vp, sz, vp]),
        "wmixb_g711_decode_device": (i, [i, vp, vp, sz, vp]),
        "wmixb_g711_bus_sum_device": (i, [vp, i, vp, vp, vp]),
        "wmixb_g711_nminus1_device": (i, [vp, i, vp, vp, vp, vp]),
        "wmixb_mix_load_device": (i, [vp, u32, u32, vp, u32, i, C.POINTER(u32), vp]),
        "wmixb_stream_state_bytes": (sz, [vp]),
        "wmixb_get_stream_state": (i, [vp, i, vp]),
        "wmixb_set_stream_state": (i, [vp, i, vp]),
        "wmixb_sync": (i, [vp]),
        "wmixb_set_tuning": (i, [vp, C.c_char_p, i]),
        "wmixb_set_default_device": (i, [i]),
        "wmixb_default_device": (i, []),
        "wmixb_tick_host_g711": (i, [vp, i, vp, vp, vp, vp, i, i]),
        "wmixb_nccl_load": (i, [C.c_char_p]),
        "wmixb_nccl_unique_id": (i, [vp]),
        "wmixb_nccl_bus_create": (i, [vp, i, i, vp, C.POINTER(vp)]),
        "wmixb_nccl_bus_destroy": (None, [vp]),
        "wmixb_nccl_bus_tick_device": (i, [vp, i, vp, vp, vp, vp]),
        "wmixb_set_default_ns_core": (i, [i]),
        "wmixb_default_ns_core": (i, []),
        "wmixb_host_alloc": (vp, [sz, i, i]),
        "wmixb_host_free": (None, [vp]),
        "wmixb_host_copy_ceiling": (i, [i, vp, vp, sz, sz, i, C.POINTER(C.c_double)]),
        "wmixb_last_error": (C.c_char_p, []),
        "wmixb_kernel_launches": (C.c_longlong, []),
        "wmixb_state_bytes_per_stream": (sz, [vp]),
        "wmixb_frame_len": (i, [vp]),
        "wmixb_selftest_fdiv": (i, [C.c_ulonglong, C.c_uint, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_ulonglong)]),
        "wmixb_ns_window": (None, [i, i, vp]),
        "wmixb_agc_gain_table": (i, [vp, i, i, i, i]),
        "wmixb_agc_analog_target": (i, [i]),
        # drop-in handle API (include/webrtc.h)
        "vad_init": (vp, [i, i, i, vp]), "vad_process": (None, [vp, vp, i]), "vad_release": (None, [vp]),
        "ns_init": (vp, [i, i, vp]), "ns_process": (None, [vp, vp, vp, i]), "ns_release": (None, [vp]),
        "agc_init": (vp, [i, i, i, i, vp]), "agc_process": (i, [vp, vp, vp, i]),
        "agc_addition": (None, [vp, C.c_uint8]), "agc_release": (None, [vp]),
        "aec_init": (vp, [i, i, i, vp]), "aec_setFrameFar": (i, [vp, vp, i]), "aec_process": (i, [vp, vp, vp, i, i]),
        "aec_process2": (i, [vp, vp, vp, vp, i, i]), "aec_release": (None, [vp]),
        # include/wmix_rtp.h
        "wmixb_rtp_write_header": (None, [vp, C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, u32, u32]),
        "wmixb_rtp_read_header": (None, [vp, vp]),
        "wmixb_rtp_unpack_device": (i, [vp, vp, i, i, i, vp, vp, vp]),
        "wmixb_rtp_pack_device": (i, [vp, i, i, vp, vp, i, vp]),
        "wmixb_mixplan_create": (i, [i, i, u32, i, i, C.POINTER(vp)]),
        "wmixb_mixplan_destroy": (None, [vp]),
        "wmixb_mixplan_out_samples": (u32, [vp]),
        "wmixb_mixplan_tables": (i, [vp, vp, vp]),
        "wmixb_mix_load_plan_device": (i, [vp, vp, u32, u32, vp, i, vp, C.POINTER(u32), vp]),
        "wmixb_load_data_host": (vp, [vp, vp, u32, C.c_uint16, C.c_uint8, C.c_uint8, vp, C.c_uint8, C.POINTER(u32)]),
        # include/wmix_zoom.h
        "wmix_len_of_out": (u32, [C.c_uint8, C.c_uint16, u32, C.c_uint8, C.c_uint16]),
        "wmix_len_of_in": (u32, [C.c_uint8, C.c_uint16, C.c_uint8, C.c_uint16, u32]),
        "wmix_pcm_zoom": (u32, [C.c_uint8, C.c_uint16, vp, u32, C.c_uint8, C.c_uint16, vp]),
        "wmixb_zoom_create": (i, [i, i, u32, i, i, i, C.POINTER(vp)]),
        "wmixb_zoom_destroy": (None, [vp]),
        "wmixb_zoom_out_bytes": (u32, [vp]),
        "wmixb_zoom_device": (i, [vp, vp, vp, i, vp]),
        "wmixb_zoom_map": (i, [vp, vp]),
        # include/g711codec.h
        "PCM2G711a": (i, [vp, vp, i, i]), "PCM2G711u": (i, [vp, vp, i, i]),
        "G711a2PCM": (i, [vp, vp, i, i]), "G711u2PCM": (i, [vp, vp, i, i]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


class WmixError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise WmixError("%s failed (%d): %s" % (what, rc, lib().wmixb_last_error().decode()))
