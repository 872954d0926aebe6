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
        L = _cabi.loaded() if (_cabi is not None and getattr(_cabi, "loaded", None) is not None) else None
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

    # ---- time-sharded calls (fdc_*_shard_*, one process per GPU; orchestration in FDC/sharded.py: ShardedActivity) ----
    def shard_measure(self, nrows, d_rows=0, stream=0, power=None):
        """Step 1: compact detection record (bytes) of this rank's rows.  d_rows: device pointer to the spectrum rows;
        host-logic contexts pass `power` (as for logic_work) instead."""
        if power is not None:
            p = np.ascontiguousarray(power, dtype=np.float32)
            n = check(self._fn("shard_measure_logic")(self._h, int(nrows), _ptr(p)), self._name)
        else:
            n = check(self._fn("shard_measure")(self._h, int(nrows), C.c_void_p(d_rows), C.c_void_p(stream) if stream else None), self._name)
        buf = C.create_string_buffer(max(int(n), 1))
        check(self._fn("shard_blob")(self._h, buf), self._name)
        return bytes(buf.raw[:int(n)])

    def shard_decide(self, nblocks_total, records):
        """Step 3: the sequential bookkeeping over the records of ALL ranks (concatenated in row order); returns the job count."""
        return int(check(self._fn("shard_decide")(self._h, int(nblocks_total), C.c_char_p(records), len(records)), self._name))

    def shard_samples(self, first_row, nrows):
        return int(check(self._fn("shard_samples")(self._h, int(first_row), int(nrows)), self._name))

    def shard_layout(self, by_channel=True):
        """device form of the call: one run per block instance for all ranks, channel by channel (fdc_*_shard_layout);
        returns the samples of the whole call"""
        return int(check(self._fn("shard_layout")(self._h, 1 if by_channel else 0), self._name))

    def shard_extract(self, first_row, nrows, d_rows=0, d_prev=0, stream=0):
        """Step 4: samples of the jobs this rank's rows emitted, in job order (complex64)."""
        out = np.empty(self.shard_samples(first_row, nrows), dtype=np.complex64)
        n = check(self._fn("shard_extract")(self._h, int(first_row), int(nrows), C.c_void_p(d_rows) if d_rows else None,
                                            C.c_void_p(d_prev) if d_prev else None, C.c_void_p(stream) if stream else None,
                                            _ptr(out) if out.size else None), self._name)
        assert int(n) == out.size
        return out

    def shard_extract_device(self, first_row, nrows, d_rows, d_prev, d_dst, stream=0):
        """Step 4, device form: the samples go to the device buffer d_dst (possibly the sink rank's, mapped over CUDA IPC)."""
        return int(check(self._fn("shard_extract_device")(self._h, int(first_row), int(nrows), C.c_void_p(d_rows), C.c_void_p(d_prev) if d_prev else None,
                                                          C.c_void_p(stream) if stream else None, C.c_void_p(d_dst)), self._name))

    def shard_assemble_device(self, d_results, nsamples, stream=0):
        """Step 5 on the sink, device form: d_results holds every rank's samples of this block in rank order."""
        check(self._fn("shard_assemble_device")(self._h, C.c_void_p(d_results), int(nsamples), C.c_void_p(stream) if stream else None), self._name)
        self._dispatch()

    def shard_assemble(self, results=None):
        """Step 5: the sink passes every rank's samples concatenated in rank order; the other ranks pass None."""
        if results is None:
            check(self._fn("shard_assemble")(self._h, None, 0), self._name)
            return
        r = np.ascontiguousarray(results, dtype=np.complex64)
        dummy = np.zeros(1, dtype=np.complex64)
        check(self._fn("shard_assemble")(self._h, _ptr(r if r.size else dummy), int(r.size)), self._name)
        self._dispatch()

    def _dispatch(self):
        if self._handler is not None:
            for m in self.messages():
                self._handler(m)

    def messages_arrays(self, clear=True, reuse=False):
        """Pending PDUs without a Python object per message: (records, data, offsets) -- `records` is a structured array with
        the PDU keys (id, finalized, part, rel_cfreq, rel_bw, blockstart, blockend, vectorstart, vectorend, nsamples), `data`
        all payloads back to back (complex64) and message k's samples are data[offsets[k]:offsets[k + 1]].  Two C calls for the
        whole batch; a busy wideband segment publishes thousands of PDUs per work() call.  reuse=True returns `data` as a view
        of a buffer that the block keeps and overwrites at the next call (no fresh pages for tens of MB per call)."""
        n = self._fn("msg_count")(self._h)
        if n <= 0:
            return np.zeros(0, dtype=_MSG_DTYPE), np.zeros(0, dtype=np.complex64), np.zeros(1, dtype=np.int64)
        recs = np.zeros(n, dtype=_MSG_DTYPE)
        check(self._fn("msg_get_all")(self._h, recs.ctypes.data_as(C.c_void_p)), self._name)
        have = (recs["data"] != 0) & (recs["nsamples"] > 0)         # host-logic contexts report counts only (data == NULL)
        offsets = np.concatenate([[0], np.cumsum(np.where(have, recs["nsamples"], 0))]).astype(np.int64)
        total = int(offsets[-1])
        if reuse:
            buf = getattr(self, "_msgbuf", None)
            if buf is None or buf.size < total:
                buf = self._msgbuf = np.empty(max(total, 1 << 16), dtype=np.complex64)
            data = buf[:total]
        else:
            data = np.empty(total, dtype=np.complex64)
        if data.size:
            check(self._fn("msg_copy_data")(self._h, _ptr(data)), self._name)
        if clear:
            self._fn("msg_clear")(self._h)
        return recs, data, offsets

    def messages(self, clear=True):
        """Pending PDUs as dicts (the `data` arrays are slices of one buffer)."""
        r, buf, cut = self.messages_arrays(clear)
        if r.size == 0:
            return []
        cut = cut.tolist()
        ids = [bytes(x).split(b"\0", 1)[0].decode() for x in r["id"].tolist()]
        cols = [r[k].tolist() for k in ("finalized", "part", "rel_bw", "rel_cfreq", "blockstart", "blockend", "vectorstart", "vectorend", "nsamples")]
        return [dict(ID=ids[k], finalized=bool(fin), part=part, rel_bw=bw, rel_cfreq=cf, blockstart=b0, blockend=b1, vectorstart=v0, vectorend=v1,
                     nsamples=ns, data=buf[cut[k]:cut[k + 1]])
                for k, (fin, part, bw, cf, b0, b1, v0, v1, ns) in enumerate(zip(*cols))]


_MSG_DTYPE = np.dtype({"names": ["id", "finalized", "part", "rel_cfreq", "rel_bw", "blockstart", "blockend", "vectorstart", "vectorend", "nsamples", "data"],
                       "formats": ["V160", "i4", "i8", "f8", "f8", "i8", "i8", "i8", "i8", "i8", "u8"],
                       "offsets": [getattr(_cabi.msg, k).offset for k in ("id", "finalized", "part", "rel_cfreq", "rel_bw", "blockstart", "blockend",
                                                                             "vectorstart", "vectorend", "nsamples", "data")],
                       "itemsize": C.sizeof(_cabi.msg)})


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
