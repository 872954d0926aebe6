"""Activity-gated blocks: FDC.PowerActivationChannel, FDC.SegmentDetection, FDC.activity_detection_channelizer_vcm.

Same constructor arguments as the reference factories (include/FDC/PowerActivationChannel.h:49,
include/FDC/SegmentDetection.h:49, include/FDC/activity_detection_channelizer_vcm.h:49).  The blocks have no stream
output; what they publish on the "msgout" message port is collected and returned by messages() as dicts with the PDU's
keys (ID, finalized, part, rel_bw, rel_cfreq, blockstart, blockend, vectorstart, vectorend) plus `data`, the c32vector
payload.  A callback registered with set_msg_handler() is invoked for every message after each work() call, which is how
the hier block forwards them (msg_connect in python/FrequencyDomainChannelizer.py:305,312).
"""
import ctypes as C
import numpy as np

from . import _cabi
from ._cabi import check, handle, lib
from .blocks import _SyncBlock, _ptr


class _MsgBlock(_SyncBlock):
    _prefix = ""
    out_itemsize = 0

    def __init__(self):
        self._handler = None

    def _fn(self, name):
        return getattr(lib(), "fdc_%s_%s" % (self._prefix, name))

    def __del__(self):
        L = _cabi.loaded() if _cabi is not None else None
        if L is not None and getattr(self, "_h", None):
            getattr(L, "fdc_%s_destroy" % self._prefix)(self._h); self._h = None

    def message_ports_out(self):
        return ["msgout"]

    def set_msg_handler(self, fn):
        self._handler = fn

    def work(self, noutput_items, input_items, output_items=None):
        """input_items[0]: noutput_items vectors of blocklen complex64 (the normalised, fft-shifted spectrum)."""
        x = np.ascontiguousarray(input_items[0])
        check(self._fn("work_host")(self._h, int(noutput_items), _ptr(x)), self._name)
        self._dispatch()
        return noutput_items

    def work_device(self, noutput_items, d_in, stream=0):
        """d_in: device pointer (int) to noutput_items spectrum rows."""
        check(self._fn("work_device")(self._h, int(noutput_items), C.c_void_p(d_in), C.c_void_p(stream) if stream else None), self._name)
        self._dispatch()
        return noutput_items

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        return self.work(x.size // self.blocklen, [x])

    def logic_work(self, nblocks, power):
        """Host-logic hook (contexts made with _logic=True only, no GPU): feed what K3 would have measured."""
        p = np.ascontiguousarray(power, dtype=np.float32)
        check(self._fn("logic_work")(self._h, int(nblocks), _ptr(p)), self._name)
        self._dispatch()
        return nblocks

    def _dispatch(self):
        if self._handler is not None:
            for m in self.messages():
                self._handler(m)

    def messages(self, clear=True):
        n = self._fn("msg_count")(self._h)
        res = []
        for k in range(n):
            m = _cabi.msg()
            check(self._fn("msg_get")(self._h, k, C.byref(m)))
            have = bool(m.data) and m.nsamples > 0         # host-logic contexts report counts only (data == NULL)
            data = np.empty(m.nsamples if have else 0, dtype=np.complex64)
            if have:
                C.memmove(data.ctypes.data, m.data, 8 * m.nsamples)
            d = dict(ID=m.id.decode(), finalized=bool(m.finalized), part=int(m.part), rel_bw=m.rel_bw, rel_cfreq=m.rel_cfreq,
                     blockstart=int(m.blockstart), blockend=int(m.blockend), vectorstart=int(m.vectorstart),
                     vectorend=int(m.vectorend), nsamples=int(m.nsamples), data=data)
            res.append(d)
        if clear:
            self._fn("msg_clear")(self._h)
        return res


class PowerActivationChannel(_MsgBlock):
    """FDC.PowerActivationChannel(blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, msg, fileoutput,
    path, verbose, ID) -- lib/PowerActivationChannel_impl.cc"""
    _name = "PowerActivationChannel"
    _prefix = "pac"

    def __init__(self, blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, msg, fileoutput, path, verbose, ID,
                 _logic=False):
        _MsgBlock.__init__(self)
        self.blocklen, self.relinvovl = int(blocklen), int(relinvovl)
        self.in_itemsize = 8 * self.blocklen
        self._h = handle((lib().fdc_pac_create_logic if _logic else lib().fdc_pac_create)(self.blocklen, float(cfreq), float(bw), self.relinvovl, float(thresh), int(maxblocks),
                                              int(deactivation_delay), int(bool(msg)), int(bool(fileoutput)), str(path).encode(),
                                              int(verbose), int(ID)), self._name)

    def state(self):
        g = (C.c_int * 12)(); f = (C.c_float * 2)()
        check(lib().fdc_pac_state(self._h, g, f))
        keys = ["extract_start", "extract_stop", "extract_width", "measure_start", "measure_stop", "deltaphase", "output_len",
                "output_ovl_offset", "active", "count", "phase", "blockcount"]
        d = {k: g[i] for i, k in enumerate(keys)}; d["thresh"] = f[0]; d["lastpower"] = f[1]
        return d

    def tables(self):
        t = np.empty((self.relinvovl, self.blocklen), dtype=np.complex64)
        check(lib().fdc_pac_tables(self._h, _ptr(t)))
        return t


class SegmentDetection(_MsgBlock):
    """FDC.SegmentDetection(ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer,
    maxblocks_to_emit, channel_deactivation_delay, messageoutput, fileoutput, path, threads, verbose)
    -- lib/SegmentDetection_impl.cc"""
    _name = "SegmentDetection"
    _prefix = "segdet"

    def __init__(self, ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer, maxblocks_to_emit,
                 channel_deactivation_delay, messageoutput, fileoutput, path, threads, verbose, _logic=False):
        _MsgBlock.__init__(self)
        self.blocklen, self.relinvovl = int(blocklen), int(relinvovl)
        self.in_itemsize = 8 * self.blocklen
        self._h = handle((lib().fdc_segdet_create_logic if _logic else lib().fdc_segdet_create)(int(ID), self.blocklen, self.relinvovl, float(seg_start), float(seg_stop), float(thresh),
                                                 float(minchandist), float(window_flank_puffer), int(maxblocks_to_emit),
                                                 int(channel_deactivation_delay), int(bool(messageoutput)), int(bool(fileoutput)),
                                                 str(path).encode(), int(bool(threads)), int(verbose)), self._name)

    def state(self):
        g = (C.c_long * 8)(); f = (C.c_float * 1)()
        check(lib().fdc_segdet_state(self._h, g, f))
        keys = ["d_start", "d_stop", "d_width", "D", "M", "blockcount", "n_active", "chan_counter"]
        d = {k: int(g[i]) for i, k in enumerate(keys)}; d["thresh"] = f[0]
        return d

    def window(self, log2w, phase):
        t = np.empty(1 << log2w, dtype=np.complex64)
        if lib().fdc_segdet_window(self._h, int(log2w), int(phase), _ptr(t)) != 0:
            raise IndexError("no such window")
        return t

    def power(self):
        p = np.empty(self.state()["M"], dtype=np.float32)
        check(lib().fdc_segdet_power(self._h, _ptr(p)))
        return p

    def active_channels(self):
        keys = ["ID", "detect_start", "detect_stop", "extract_start", "extract_stop", "extract_width", "ovlskip", "outputsamples",
                "count", "phase", "phaseincrement", "inactive", "part", "ndata"]
        res = []
        for k in range(self.state()["n_active"]):
            a = (C.c_int * 14)()
            check(lib().fdc_segdet_active(self._h, k, a))
            res.append({kk: a[i] for i, kk in enumerate(keys)})
        return res


class activity_detection_channelizer_vcm(_MsgBlock):
    """FDC.activity_detection_channelizer_vcm(blocklen, segments, thresh, relinvovl, maxblocks, message, fileoutput, path,
    threads, minchandist, channel_deactivation_delay, window_flank_puffer, verbose)
    -- lib/activity_detection_channelizer_vcm_impl.cc"""
    _name = "activity_detection_channelizer_vcm"
    _prefix = "actdet"

    def __init__(self, blocklen, segments, thresh, relinvovl, maxblocks, message, fileoutput, path, threads, minchandist,
                 channel_deactivation_delay, window_flank_puffer, verbose, _logic=False):
        _MsgBlock.__init__(self)
        self.blocklen, self.relinvovl = int(blocklen), int(relinvovl)
        self.in_itemsize = 8 * self.blocklen
        for s in segments:
            if len(s) != 2:
                raise _cabi.FDCError("Segment is incorrect. must be of size 2 with each member in (0,1), with v[0]<v[1]. ")
        segs = np.ascontiguousarray(np.asarray(segments, dtype=np.float32).reshape(-1, 2))
        self._h = handle((lib().fdc_actdet_create_logic if _logic else lib().fdc_actdet_create)(self.blocklen, _ptr(segs), segs.shape[0], float(thresh), self.relinvovl, int(maxblocks),
                                                 int(bool(message)), int(bool(fileoutput)), str(path).encode(), int(bool(threads)),
                                                 float(minchandist), int(channel_deactivation_delay), float(window_flank_puffer),
                                                 int(verbose)), self._name)

    def segments(self):
        keys = ["ID", "start", "stop", "width", "D", "M", "n_active"]
        res = []
        for k in range(lib().fdc_actdet_nsegments(self._h)):
            a = (C.c_int * 7)()
            check(lib().fdc_actdet_segment(self._h, k, a))
            res.append({kk: a[i] for i, kk in enumerate(keys)})
        return res

    def power(self, seg):
        p = np.empty(self.segments()[seg]["M"], dtype=np.float32)
        check(lib().fdc_actdet_power(self._h, int(seg), _ptr(p)))
        return p
