"""ctypes binding of libfdc_b200.so (include/fdc_cabi.h).

This is the only place the Python host side touches native code.  There is no CPU fallback:
if the library is missing, or no CUDA device is usable, the block constructors raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDC_LIB_PATH") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libfdc_b200.so"))     # FDC_LIB_PATH: an experimental build (tools/)
_lib = None


class FDCError(RuntimeError):
    """The reference's std::invalid_argument reaches Python as RuntimeError through SWIG; same type here."""


class chan_desc(C.Structure):
    _fields_ = [("f", C.c_int), ("l", C.c_int), ("lout", C.c_int), ("shift", C.c_int), ("gain", C.c_float),
                ("table", C.c_void_p)]


class msg(C.Structure):
    _fields_ = [("id", C.c_char * 160), ("finalized", C.c_int), ("part", C.c_long), ("rel_cfreq", C.c_double),
                ("rel_bw", C.c_double), ("blockstart", C.c_long), ("blockend", C.c_long), ("vectorstart", C.c_long),
                ("vectorend", C.c_long), ("nsamples", C.c_long), ("data", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/fdc_cabi.h declares
_vp, _i, _l, _f, _d, _cp = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double, C.c_char_p
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
SIGNATURES = {
    "fdc_api_version": (_i, []),
    "fdc_last_error": (_cp, []),
    "fdc_device_count": (_i, []),
    "fdc_set_device": (_i, [_i]),
    "fdc_host_alloc": (_vp, [C.c_size_t]),
    "fdc_host_free": (None, [_vp]),
    "fdc_copy_threads": (_i, []),
    "fdc_host_register": (_i, [_vp, C.c_size_t]),
    "fdc_host_unregister": (_i, [_vp]),
    "fdc_host_evict": (None, [_vp, C.c_size_t]),
    "fdc_launch_count": (C.c_ulonglong, []),
    "fdc_dev_alloc": (_vp, [C.c_size_t]),
    "fdc_dev_free": (None, [_vp]),
    "fdc_memcpy_h2d": (_i, [_vp, _vp, C.c_size_t]),
    "fdc_memcpy_d2h": (_i, [_vp, _vp, C.c_size_t]),
    "fdc_device_synchronize": (_i, []),
    "fdc_opt_channelparams": (_i, [_i, _i, _d, _d, _ip, _ip, _ip, _dp, _dp]),
    "fdc_psw_build_tables": (_i, [_i, _i, _f, _f, _i, _vp]),
    "fdc_chan_create": (_vp, [_i, _i, _i, _i, _vp]),
    "fdc_chan_destroy": (None, [_vp]),
    "fdc_chan_hop": (_i, [_vp]),
    "fdc_chan_blockcount": (_l, [_vp]),
    "fdc_chan_reset": (_i, [_vp]),
    "fdc_chan_seek": (_i, [_vp, _l]),
    "fdc_chan_set_history": (_i, [_vp, _vp]),
    "fdc_chan_work_host": (_i, [_vp, _vp, _l, _vp, _vp]),
    "fdc_chan_work_device": (_i, [_vp, _vp, _l, _vp, _vp, _vp]),
    "fdc_chan_work_device_slab": (_i, [_vp, _vp, _l, _vp, _l, _l, _vp, _vp]),
    "fdc_ipc_export": (_i, [_vp, _vp]),
    "fdc_ipc_open": (_vp, [_vp]),
    "fdc_ipc_close": (_i, [_vp]),
    "fdc_chan_work_spectrum_device": (_i, [_vp, _vp, _l, _vp, _vp, _vp]),
    "fdc_chan_sync": (_i, [_vp]),
    "fdc_chan_set_profiling": (_i, [_vp, _i]),
    "fdc_chan_get_profile": (_i, [_vp, _dp, _dp, C.POINTER(C.c_long)]),
    "fdc_chan_set_sinks": (_i, [_vp, _i, _vp, _vp, _i]),
    "fdc_chan_work_device_sinks": (_i, [_vp, _vp, _l, _l, _l, _vp]),
    "fdc_chan_is_fused": (_i, [_vp]),
    "fdc_chan_chunk_blocks": (_i, [_vp]),
    "fdc_chan_set_chunk_blocks": (_i, [_vp, _i]),
    "fdc_overlap_save_create": (_vp, [_i, _i, _i]),
    "fdc_overlap_save_work": (_i, [_vp, _i, _vp, _vp]),
    "fdc_overlap_save_destroy": (None, [_vp]),
    "fdc_vector_cut_create": (_vp, [_i, _i, _i, _i]),
    "fdc_vector_cut_work": (_i, [_vp, _i, _vp, _vp]),
    "fdc_vector_cut_destroy": (None, [_vp]),
    "fdc_psw_create": (_vp, [_i, _i, _i, _f, _f, _i]),
    "fdc_psw_work": (_i, [_vp, _i, _vp, _vp]),
    "fdc_psw_state": (_i, [_vp, _ip, _ip, _ip, _ip]),
    "fdc_psw_tables": (_i, [_vp, _vp]),
    "fdc_psw_destroy": (None, [_vp]),
    "fdc_fft_create": (_vp, [_i, _i, _i]),
    "fdc_fft_work": (_i, [_vp, _l, _vp, _vp]),
    "fdc_fft_destroy": (None, [_vp]),
    "fdc_waterfall_create": (_vp, [_i, _i, _i]),
    "fdc_waterfall_work_host": (_i, [_vp, _i, _vp, _vp]),
    "fdc_waterfall_work_device": (_i, [_vp, _i, _vp, _vp, _vp]),
    "fdc_waterfall_destroy": (None, [_vp]),
    "fdc_pac_create": (_vp, [_i, _f, _f, _i, _f, _i, _i, _i, _i, _cp, _i, _i]),
    "fdc_pac_work_host": (_i, [_vp, _i, _vp]),
    "fdc_pac_work_device": (_i, [_vp, _i, _vp, _vp]),
    "fdc_pac_state": (_i, [_vp, _vp, _vp]),
    "fdc_pac_tables": (_i, [_vp, _vp]),
    "fdc_pac_msg_count": (_i, [_vp]),
    "fdc_pac_msg_get": (_i, [_vp, _i, _vp]),
    "fdc_pac_msg_get_all": (_i, [_vp, _vp]),
    "fdc_pac_msg_copy_data": (_l, [_vp, _vp]),
    "fdc_pac_msg_clear": (None, [_vp]),
    "fdc_pac_destroy": (None, [_vp]),
    "fdc_segdet_create": (_vp, [_i, _i, _i, _f, _f, _f, _f, _f, _i, _i, _i, _i, _cp, _i, _i]),
    "fdc_segdet_work_host": (_i, [_vp, _i, _vp]),
    "fdc_segdet_work_device": (_i, [_vp, _i, _vp, _vp]),
    "fdc_segdet_state": (_i, [_vp, _vp, _vp]),
    "fdc_segdet_window": (_i, [_vp, _i, _i, _vp]),
    "fdc_segdet_power": (_i, [_vp, _vp]),
    "fdc_segdet_active": (_i, [_vp, _i, _vp]),
    "fdc_segdet_msg_count": (_i, [_vp]),
    "fdc_segdet_msg_get": (_i, [_vp, _i, _vp]),
    "fdc_segdet_msg_get_all": (_i, [_vp, _vp]),
    "fdc_segdet_msg_copy_data": (_l, [_vp, _vp]),
    "fdc_segdet_msg_clear": (None, [_vp]),
    "fdc_segdet_destroy": (None, [_vp]),
    "fdc_actdet_create": (_vp, [_i, _vp, _i, _f, _i, _i, _i, _i, _cp, _i, _f, _i, _d, _i]),
    "fdc_actdet_work_host": (_i, [_vp, _i, _vp]),
    "fdc_actdet_work_device": (_i, [_vp, _i, _vp, _vp]),
    "fdc_actdet_nsegments": (_i, [_vp]),
    "fdc_actdet_segment": (_i, [_vp, _i, _vp]),
    "fdc_actdet_power": (_i, [_vp, _i, _vp]),
    "fdc_actdet_msg_count": (_i, [_vp]),
    "fdc_actdet_msg_get": (_i, [_vp, _i, _vp]),
    "fdc_actdet_msg_get_all": (_i, [_vp, _vp]),
    "fdc_actdet_msg_copy_data": (_l, [_vp, _vp]),
    "fdc_actdet_msg_clear": (None, [_vp]),
    "fdc_actdet_destroy": (None, [_vp]),
    "fdc_pac_create_logic": (_vp, [_i, _f, _f, _i, _f, _i, _i, _i, _i, _cp, _i, _i]),
    "fdc_pac_logic_work": (_i, [_vp, _i, _vp]),
    "fdc_segdet_create_logic": (_vp, [_i, _i, _i, _f, _f, _f, _f, _f, _i, _i, _i, _i, _cp, _i, _i]),
    "fdc_segdet_logic_work": (_i, [_vp, _i, _vp]),
    "fdc_actdet_create_logic": (_vp, [_i, _vp, _i, _f, _i, _i, _i, _i, _cp, _i, _f, _i, _d, _i]),
    "fdc_actdet_logic_work": (_i, [_vp, _i, _vp]),
    "fdc_pac_shard_measure": (_l, [_vp, _i, _vp, _vp]),
    "fdc_pac_shard_measure_logic": (_l, [_vp, _i, _vp]),
    "fdc_pac_shard_blob": (_i, [_vp, _vp]),
    "fdc_pac_shard_decide": (_l, [_vp, _i, _vp, _l]),
    "fdc_pac_shard_samples": (_l, [_vp, _i, _i]),
    "fdc_pac_shard_layout": (_l, [_vp, _i]),
    "fdc_pac_shard_extract": (_l, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "fdc_pac_shard_assemble": (_i, [_vp, _vp, _l]),
    "fdc_pac_shard_extract_device": (_l, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "fdc_pac_shard_assemble_device": (_i, [_vp, _vp, _l, _vp]),
    "fdc_segdet_shard_measure": (_l, [_vp, _i, _vp, _vp]),
    "fdc_segdet_shard_measure_logic": (_l, [_vp, _i, _vp]),
    "fdc_segdet_shard_blob": (_i, [_vp, _vp]),
    "fdc_segdet_shard_decide": (_l, [_vp, _i, _vp, _l]),
    "fdc_segdet_shard_samples": (_l, [_vp, _i, _i]),
    "fdc_segdet_shard_layout": (_l, [_vp, _i]),
    "fdc_segdet_shard_extract": (_l, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "fdc_segdet_shard_assemble": (_i, [_vp, _vp, _l]),
    "fdc_segdet_shard_extract_device": (_l, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "fdc_segdet_shard_assemble_device": (_i, [_vp, _vp, _l, _vp]),
    "fdc_actdet_shard_measure": (_l, [_vp, _i, _vp, _vp]),
    "fdc_actdet_shard_measure_logic": (_l, [_vp, _i, _vp]),
    "fdc_actdet_shard_blob": (_i, [_vp, _vp]),
    "fdc_actdet_shard_decide": (_l, [_vp, _i, _vp, _l]),
    "fdc_actdet_shard_samples": (_l, [_vp, _i, _i]),
    "fdc_actdet_shard_layout": (_l, [_vp, _i]),
    "fdc_actdet_shard_extract": (_l, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "fdc_actdet_shard_assemble": (_i, [_vp, _vp, _l]),
    "fdc_actdet_shard_extract_device": (_l, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "fdc_actdet_shard_assemble_device": (_i, [_vp, _vp, _l, _vp]),
}


def lib():
    """Load libfdc_b200.so (once).  Raises FDCError when it has not been built."""
    global _lib
    if _lib is None:
        if os is None:                 # interpreter shutdown: module globals are being torn down, nothing to load any more
            raise FDCError("interpreter is shutting down")
        if not os.path.exists(LIB_PATH):
            raise FDCError("libfdc_b200.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C gr-fdc_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name, None)
            if fn is None:
                raise FDCError("libfdc_b200.so does not export %s (stale build?)" % name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def loaded():
    """The library if it has been loaded already, else None (destructors use this: they must not load or raise)."""
    return _lib


def last_error():
    return lib().fdc_last_error().decode(errors="replace")


def check(status, what=""):
    if status is None or (isinstance(status, int) and status < 0):
        raise FDCError((what + ": " if what else "") + last_error())
    return status


def handle(ptr, what=""):
    if not ptr:
        raise FDCError((what + ": " if what else "") + last_error())
    return C.c_void_p(ptr)
