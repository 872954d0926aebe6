"""Oracle (TEST INFRASTRUCTURE ONLY): ctypes binding of oracle/_ref/libfdc_ref.so, i.e. the
UNMODIFIED gr-FDC C++ blocks (/root/reference/lib/*_impl.cc) compiled against oracle/shim.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (gr-fdc_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libfdc_ref.so")
_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        vp, i, f, d, cp = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_char_p
        L.ref_last_error.restype = cp
        L.ref_overlap_save_make.restype = vp; L.ref_overlap_save_make.argtypes = [i, i, i]
        L.ref_vector_cut_make.restype = vp; L.ref_vector_cut_make.argtypes = [i, i, i, i]
        L.ref_psw_make.restype = vp; L.ref_psw_make.argtypes = [i, i, i, f, f, i]
        L.ref_pac_make.restype = vp; L.ref_pac_make.argtypes = [i, f, f, i, f, i, i, i, i, cp, i, i]
        L.ref_segdet_make.restype = vp
        L.ref_segdet_make.argtypes = [i, i, i, f, f, f, f, f, i, i, i, i, cp, i, i]
        L.ref_actdet_make.restype = vp
        L.ref_actdet_make.argtypes = [i, vp, i, f, i, i, i, i, cp, i, f, i, d, i]
        L.ref_work.restype = i; L.ref_work.argtypes = [vp, i, vp, vp]
        L.ref_free.argtypes = [vp]
        L.ref_name.restype = cp; L.ref_name.argtypes = [vp]
        L.ref_itemsizes.argtypes = [vp, vp, vp]
        L.ref_msg_count.restype = i; L.ref_msg_count.argtypes = [vp]
        L.ref_msg_clear.argtypes = [vp]
        L.ref_msg_meta.restype = i; L.ref_msg_meta.argtypes = [vp, i, vp, i, vp, vp]
        L.ref_msg_data.restype = i; L.ref_msg_data.argtypes = [vp, i, vp]
        L.ref_msg_keys.restype = i; L.ref_msg_keys.argtypes = [vp, i, vp, i]
        L.ref_psw_tables.argtypes = [vp, vp]; L.ref_psw_state.argtypes = [vp, vp]
        L.ref_pac_state.argtypes = [vp, vp, vp]; L.ref_pac_tables.argtypes = [vp, vp]
        L.ref_segdet_state.argtypes = [vp, vp, vp]; L.ref_segdet_window.argtypes = [vp, i, i, vp]
        L.ref_segdet_power.argtypes = [vp, vp]; L.ref_segdet_active.argtypes = [vp, i, vp]
        L.ref_actdet_segment.argtypes = [vp, i, vp]; L.ref_actdet_nsegments.argtypes = [vp]
        L.ref_actdet_power.argtypes = [vp, i, vp]
        L.ref_fft_vcc.restype = i; L.ref_fft_vcc.argtypes = [i, i, i, C.c_long, vp, vp]
        L.ref_multiply_const_cc.argtypes = [f, C.c_long, vp, vp]
        L.ref_chain_make.restype = vp; L.ref_chain_make.argtypes = [i, i, i, vp, vp, vp, vp, vp, i]
        L.ref_chain_free.argtypes = [vp]
        L.ref_chain_run.restype = i; L.ref_chain_run.argtypes = [vp, vp, C.c_long, vp, vp, i]
        L.fdc_shim_set_fft_mode.argtypes = [i]
        L.fdc_shim_get_fft_mode.restype = i
        _lib = L
    return _lib


def set_fft_mode(mode):
    """0 = fp64-accurate (parity anchor), 1 = fp32 Stockham (timing)."""
    lib().fdc_shim_set_fft_mode(int(mode))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefError(RuntimeError):
    """SWIG maps the blocks' std::invalid_argument to RuntimeError; same here."""


class Block:
    """One reference sync_block instance.  work() mirrors gr::sync_block::work (1:1 items)."""

    def __init__(self, handle):
        if not handle:
            raise RefError(lib().ref_last_error().decode())
        self.h = C.c_void_p(handle)
        a, b = C.c_int(), C.c_int()
        lib().ref_itemsizes(self.h, C.byref(a), C.byref(b))
        self.in_itemsize, self.out_itemsize = a.value, b.value
        self.name = lib().ref_name(self.h).decode()

    def __del__(self):
        try:
            if self.h:
                lib().ref_free(self.h); self.h = None
        except Exception:
            pass

    def work(self, inp, nitems=None):
        inp = np.ascontiguousarray(inp)
        if nitems is None:
            nitems = inp.nbytes // self.in_itemsize
        assert inp.nbytes >= nitems * self.in_itemsize
        out = np.empty(max(nitems * self.out_itemsize, 1), dtype=np.uint8)
        r = lib().ref_work(self.h, int(nitems), _ptr(inp), _ptr(out))
        if r != nitems:
            raise RefError(lib().ref_last_error().decode())
        return out[: nitems * self.out_itemsize]

    # ---- captured PDUs ----
    def message_keys(self):
        """dict keys of every captured PDU in the order the block added them"""
        L = lib(); res = []
        for k in range(L.ref_msg_count(self.h)):
            buf = C.create_string_buffer(512)
            L.ref_msg_keys(self.h, k, buf, 512)
            res.append(buf.value.decode().split(","))
        return res

    def messages(self, clear=True):
        L = lib(); n = L.ref_msg_count(self.h); res = []
        for k in range(n):
            idb = C.create_string_buffer(256); ints = (C.c_long * 7)(); dbl = (C.c_double * 2)()
            L.ref_msg_meta(self.h, k, idb, 256, ints, dbl)
            data = np.empty(ints[6], dtype=np.complex64)
            if ints[6]:
                L.ref_msg_data(self.h, k, _ptr(data))
            res.append(dict(ID=idb.value.decode(), finalized=bool(ints[0]), part=int(ints[1]),
                            blockstart=int(ints[2]), blockend=int(ints[3]), vectorstart=int(ints[4]),
                            vectorend=int(ints[5]), rel_bw=dbl[0], rel_cfreq=dbl[1], data=data))
        if clear:
            L.ref_msg_clear(self.h)
        return res


def overlap_save(itemsize, outputlen, overlaplen):
    return Block(lib().ref_overlap_save_make(itemsize, outputlen, overlaplen))


def vector_cut_vxx(itemsize, veclen, offset, blocklen):
    return Block(lib().ref_vector_cut_make(itemsize, veclen, offset, blocklen))


class phase_shifting_windowing_vcc(Block):
    def __init__(self, blocklen, numphasestates, shifts, passbw, stopbw, windowtype):
        super().__init__(lib().ref_psw_make(blocklen, numphasestates, shifts, passbw, stopbw, windowtype))

    def state(self):
        st = (C.c_int * 4)(); lib().ref_psw_state(self.h, st)
        return dict(blocksize=st[0], relinvovl=st[1], counter=st[2], shift=st[3])

    def tables(self):
        st = self.state()
        t = np.empty((st["relinvovl"], st["blocksize"]), dtype=np.complex64)
        lib().ref_psw_tables(self.h, _ptr(t)); return t


class PowerActivationChannel(Block):
    def __init__(self, blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, msg, fileoutput,
                 path, verbose, ID):
        self.blocklen, self.relinvovl = blocklen, relinvovl
        super().__init__(lib().ref_pac_make(blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay,
                                            int(msg), int(fileoutput), str(path).encode(), verbose, ID))

    def state(self):
        g = (C.c_int * 12)(); f = (C.c_float * 2)(); lib().ref_pac_state(self.h, g, f)
        keys = ["extract_start", "extract_stop", "extract_width", "measure_start", "measure_stop", "deltaphase",
                "output_len", "output_ovl_offset", "active", "count", "phase", "blockcount"]
        d = {k: g[i] for i, k in enumerate(keys)}; d["thresh"] = f[0]; d["lastpower"] = f[1]; return d

    def tables(self):
        t = np.empty((self.relinvovl, self.blocklen), dtype=np.complex64)
        lib().ref_pac_tables(self.h, _ptr(t)); return t


class SegmentDetection(Block):
    def __init__(self, ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer,
                 maxblocks_to_emit, channel_deactivation_delay, messageoutput, fileoutput, path, threads, verbose):
        self.blocklen, self.relinvovl = blocklen, relinvovl
        super().__init__(lib().ref_segdet_make(ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist,
                                               window_flank_puffer, maxblocks_to_emit, channel_deactivation_delay,
                                               int(messageoutput), int(fileoutput), str(path).encode(), int(threads),
                                               verbose))

    def state(self):
        g = (C.c_long * 8)(); f = (C.c_float * 1)(); lib().ref_segdet_state(self.h, g, f)
        keys = ["d_start", "d_stop", "d_width", "D", "M", "blockcount", "n_active", "chan_counter"]
        d = {k: int(g[i]) for i, k in enumerate(keys)}; d["thresh"] = f[0]; return d

    def window(self, log2w, phase):
        t = np.empty(1 << log2w, dtype=np.complex64)
        if lib().ref_segdet_window(self.h, log2w, phase, _ptr(t)) != 0:
            raise IndexError("no such window")
        return t

    def power(self):
        p = np.empty(self.state()["M"], dtype=np.float32); lib().ref_segdet_power(self.h, _ptr(p)); return p

    def active_channels(self):
        keys = ["ID", "detect_start", "detect_stop", "extract_start", "extract_stop", "extract_width", "ovlskip",
                "outputsamples", "count", "phase", "phaseincrement", "inactive", "part", "ndata"]
        res = []
        for k in range(self.state()["n_active"]):
            a = (C.c_int * 14)(); lib().ref_segdet_active(self.h, k, a)
            res.append({kk: a[i] for i, kk in enumerate(keys)})
        return res


class activity_detection_channelizer_vcm(Block):
    def __init__(self, blocklen, segments, thresh, relinvovl, maxblocks, message, fileoutput, path, threads,
                 minchandist, channel_deactivation_delay, window_flank_puffer, verbose):
        segs = np.ascontiguousarray(np.asarray(segments, dtype=np.float32).reshape(-1, 2))
        super().__init__(lib().ref_actdet_make(blocklen, _ptr(segs), segs.shape[0], thresh, relinvovl, maxblocks,
                                               int(message), int(fileoutput), str(path).encode(), int(threads),
                                               minchandist, channel_deactivation_delay, window_flank_puffer, verbose))

    def segments(self):
        keys = ["ID", "start", "stop", "width", "D", "M", "n_active"]; res = []
        for k in range(lib().ref_actdet_nsegments(self.h)):
            a = (C.c_int * 7)(); lib().ref_actdet_segment(self.h, k, a)
            res.append({kk: a[i] for i, kk in enumerate(keys)})
        return res

    def power(self, seg):
        p = np.empty(self.segments()[seg]["M"], dtype=np.float32); lib().ref_actdet_power(self.h, seg, _ptr(p)); return p


def fft_vcc(x, n, forward, shift):
    """Restated third-party gr-fft fft_vcc with an all-ones window (see ref_driver.cc)."""
    x = np.ascontiguousarray(x, dtype=np.complex64); nvec = x.size // n
    out = np.empty(nvec * n, dtype=np.complex64)
    if lib().ref_fft_vcc(n, int(forward), int(shift), nvec, _ptr(x), _ptr(out)) != 0:
        raise RefError(lib().ref_last_error().decode())
    return out


class Chain:
    """The hier block's throughput flowgraph built from the reference blocks (ref_driver.cc)."""

    def __init__(self, N, R, params, windowtype):
        """params: list of (f, l, lout, passbw, stopbw) as returned by get_opt_channelparams."""
        self.N, self.R, self.hop = N, R, N - N // R
        self.f = np.array([p[0] for p in params], dtype=np.int32)
        self.l = np.array([p[1] for p in params], dtype=np.int32)
        self.lout = np.array([p[2] for p in params], dtype=np.int32)
        pbw = np.array([p[3] for p in params], dtype=np.float32)
        sbw = np.array([p[4] for p in params], dtype=np.float32)
        h = lib().ref_chain_make(N, R, len(params), _ptr(self.f), _ptr(self.l), _ptr(self.lout), _ptr(pbw), _ptr(sbw),
                                 windowtype)
        if not h:
            raise RefError(lib().ref_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                lib().ref_chain_free(self.h); self.h = None
        except Exception:
            pass

    def run(self, x, nthreads=1, want_spectrum=False, want_outputs=True):
        x = np.ascontiguousarray(x, dtype=np.complex64); nblocks = x.size // self.hop
        outs = [np.empty(nblocks * int(lo), dtype=np.complex64) for lo in self.lout] if want_outputs else []
        arr = (C.c_void_p * len(self.lout))(*[o.ctypes.data for o in outs]) if want_outputs else None
        spec = np.empty(nblocks * self.N, dtype=np.complex64) if want_spectrum else None
        r = lib().ref_chain_run(self.h, _ptr(x), nblocks, arr, _ptr(spec) if want_spectrum else None, nthreads)
        if r != 0:
            raise RefError(lib().ref_last_error().decode())
        return outs, spec
