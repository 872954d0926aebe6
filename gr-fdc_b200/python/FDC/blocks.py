"""Host-side mirrors of the gr-FDC blocks, same names and constructor arguments as the reference's
`FDC.<block>(...)` SWIG factories (include/FDC/*.h:49), running on libfdc_b200.so.

GNU Radio is not required: every block has the scheduler-facing

    work(noutput_items, input_items, output_items) -> noutput_items

of gr::sync_block (input_items / output_items are lists of numpy arrays, one item = one vector), plus a
convenience `process(array) -> array`.  State carried across work() calls (overlap history, phase counter,
activity state machines) makes the result independent of how the stream is chunked, as in the reference.
"""
import ctypes as C
import numpy as np

from . import _cabi
from ._cabi import FDCError, check, handle, lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as_bytes(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


class _SyncBlock(object):
    """Minimal stand-in for the gr::sync_block surface the flowgraph uses."""
    _name = "sync_block"
    in_itemsize = 0
    out_itemsize = 0

    def name(self):
        return self._name

    def input_signature(self):
        return (1, 1, self.in_itemsize)

    def output_signature(self):
        return (1, 1, self.out_itemsize) if self.out_itemsize else (0, 0, 0)


class overlap_save(_SyncBlock):
    """FDC.overlap_save(itemsize, outputlen, overlaplen) -- lib/overlap_save_impl.cc"""
    _name = "overlap_save"

    def __init__(self, itemsize, outputlen, overlaplen):
        self.itemsize, self.outputlen, self.overlaplen = int(itemsize), int(outputlen), int(overlaplen)
        self.in_itemsize = self.itemsize * (self.outputlen - self.overlaplen)
        self.out_itemsize = self.itemsize * self.outputlen
        self._h = handle(lib().fdc_overlap_save_create(self.itemsize, self.outputlen, self.overlaplen), "overlap_save")

    def __del__(self):
        L = _cabi.loaded() if (_cabi is not None and getattr(_cabi, "loaded", None) is not None) else None     # module globals are None at interpreter shutdown
        if L is not None and getattr(self, "_h", None):
            L.fdc_overlap_save_destroy(self._h); self._h = None

    def work(self, noutput_items, input_items, output_items):
        check(lib().fdc_overlap_save_work(self._h, int(noutput_items), _ptr(input_items[0]), _ptr(output_items[0])), "overlap_save")
        return noutput_items

    def process(self, x):
        raw = _as_bytes(x); n = raw.size // self.in_itemsize
        out = np.empty(n * self.out_itemsize, dtype=np.uint8)
        self.work(n, [raw], [out])
        return out.view(np.asarray(x).dtype) if out.size % np.asarray(x).dtype.itemsize == 0 else out


class vector_cut_vxx(_SyncBlock):
    """FDC.vector_cut_vxx(itemsize, veclen, offset, blocklen) -- lib/vector_cut_vxx_impl.cc"""
    _name = "vector_cut_vxx"

    def __init__(self, itemsize, veclen, offset, blocklen):
        self.itemsize, self.veclen, self.offset, self.blocklen = int(itemsize), int(veclen), int(offset), int(blocklen)
        self.in_itemsize = self.itemsize * self.veclen
        self.out_itemsize = self.itemsize * self.blocklen
        self._h = handle(lib().fdc_vector_cut_create(self.itemsize, self.veclen, self.offset, self.blocklen), "vector_cut_vxx")

    def __del__(self):
        L = _cabi.loaded() if (_cabi is not None and getattr(_cabi, "loaded", None) is not None) else None     # module globals are None at interpreter shutdown
        if L is not None and getattr(self, "_h", None):
            L.fdc_vector_cut_destroy(self._h); self._h = None

    def work(self, noutput_items, input_items, output_items):
        check(lib().fdc_vector_cut_work(self._h, int(noutput_items), _ptr(input_items[0]), _ptr(output_items[0])), "vector_cut_vxx")
        return noutput_items

    def process(self, x):
        raw = _as_bytes(x); n = raw.size // self.in_itemsize
        out = np.empty(n * self.out_itemsize, dtype=np.uint8)
        self.work(n, [raw], [out])
        return out.view(np.asarray(x).dtype) if out.size % np.asarray(x).dtype.itemsize == 0 else out


class phase_shifting_windowing_vcc(_SyncBlock):
    """FDC.phase_shifting_windowing_vcc(blocklen, numphasestates, shifts, passbw, stopbw, windowtype)
    -- lib/phase_shifting_windowing_vcc_impl.cc"""
    _name = "phase_shifting_windowing_vcc"

    def __init__(self, blocklen, numphasestates, shifts, passbw, stopbw, windowtype):
        self.blocklen = int(blocklen)
        self.in_itemsize = self.out_itemsize = 8 * self.blocklen
        self._h = handle(lib().fdc_psw_create(self.blocklen, int(numphasestates), int(shifts), float(passbw), float(stopbw),
                                              int(windowtype)), "phase_shifting_windowing_vcc")

    def __del__(self):
        L = _cabi.loaded() if (_cabi is not None and getattr(_cabi, "loaded", None) is not None) else None     # module globals are None at interpreter shutdown
        if L is not None and getattr(self, "_h", None):
            L.fdc_psw_destroy(self._h); self._h = None

    def work(self, noutput_items, input_items, output_items):
        check(lib().fdc_psw_work(self._h, int(noutput_items), _ptr(input_items[0]), _ptr(output_items[0])), "psw")
        return noutput_items

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64); n = x.size // self.blocklen
        out = np.empty(n * self.blocklen, dtype=np.complex64)
        self.work(n, [x], [out])
        return out

    def state(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().fdc_psw_state(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(blocksize=a.value, relinvovl=b.value, counter=c.value, shift=d.value)

    def tables(self):
        st = self.state()
        t = np.empty((st["relinvovl"], st["blocksize"]), dtype=np.complex64)
        check(lib().fdc_psw_tables(self._h, _ptr(t)))
        return t


class fft_vcc(_SyncBlock):
    """gr::fft::fft_vcc(fft_size, forward, rectangular window, shift) as used at
    python/FrequencyDomainChannelizer.py:206,228 (third-party stage; window must be all ones)."""
    _name = "fft_vcc"

    def __init__(self, fft_size, forward, window=None, shift=False, nthreads=1):
        if window is not None and len(window) and not np.all(np.asarray(window) == 1.0):
            raise FDCError("fft_vcc: only the rectangular window the hier block uses is supported")
        self.n = int(fft_size)
        self.in_itemsize = self.out_itemsize = 8 * self.n
        self._h = handle(lib().fdc_fft_create(self.n, int(bool(forward)), int(bool(shift))), "fft_vcc")

    def __del__(self):
        L = _cabi.loaded() if (_cabi is not None and getattr(_cabi, "loaded", None) is not None) else None     # module globals are None at interpreter shutdown
        if L is not None and getattr(self, "_h", None):
            L.fdc_fft_destroy(self._h); self._h = None

    def work(self, noutput_items, input_items, output_items):
        check(lib().fdc_fft_work(self._h, int(noutput_items), _ptr(input_items[0]), _ptr(output_items[0])), "fft_vcc")
        return noutput_items

    def process(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64); n = x.size // self.n
        out = np.empty(n * self.n, dtype=np.complex64)
        self.work(n, [x], [out])
        return out


def opt_channelparams(blocksize, relinvovl, freq, bw):
    """FrequencyDomainChannelizer.get_opt_channelparams (python/FrequencyDomainChannelizer.py:322-345),
    computed by the native library (the same routine the C++ host side uses)."""
    f, l, lo = C.c_int(), C.c_int(), C.c_int(); pb, sb = C.c_double(), C.c_double()
    check(lib().fdc_opt_channelparams(int(blocksize), int(relinvovl), float(freq), float(bw), C.byref(f), C.byref(l), C.byref(lo),
                                      C.byref(pb), C.byref(sb)), "get_opt_channelparams")
    return f.value, l.value, lo.value, pb.value, sb.value


def psw_tables(blocklen, numphasestates, passbw, stopbw, windowtype):
    t = np.empty((int(numphasestates), int(blocklen)), dtype=np.complex64)
    check(lib().fdc_psw_build_tables(int(blocklen), int(numphasestates), float(passbw), float(stopbw), int(windowtype), _ptr(t)),
          "phase_shifting_windowing_vcc")
    return t


class Channelizer(object):
    """The fused throughput path (fdc_chan_*): overlap-save -> forward FFT -> per-channel extract.

    channels: list of dicts/tuples (f, l, lout, shift, gain, table[nphase, l])."""

    def __init__(self, N, ovl, nphase, channels):
        self.N, self.ovl, self.hop, self.nphase = int(N), int(ovl), int(N) - int(ovl), int(nphase)
        self.f = [int(c[0]) for c in channels]; self.l = [int(c[1]) for c in channels]
        self.lout = [int(c[2]) for c in channels]
        self._tables = [np.ascontiguousarray(c[5], dtype=np.complex64).reshape(self.nphase, int(c[1])) for c in channels]
        arr = (_cabi.chan_desc * max(len(channels), 1))()
        for i, c in enumerate(channels):
            arr[i].f, arr[i].l, arr[i].lout, arr[i].shift = int(c[0]), int(c[1]), int(c[2]), int(c[3])
            arr[i].gain = float(c[4]); arr[i].table = self._tables[i].ctypes.data
        self.nchan = len(channels)
        self.lout_prefix = np.concatenate([[0], np.cumsum(self.lout)]).astype(np.int64)
        self._h = handle(lib().fdc_chan_create(self.N, self.ovl, self.nphase, self.nchan, C.cast(arr, C.c_void_p)), "fdc_chan_create")

    def __del__(self):
        L = _cabi.loaded() if (_cabi is not None and getattr(_cabi, "loaded", None) is not None) else None     # module globals are None at interpreter shutdown
        if L is not None and getattr(self, "_h", None):
            L.fdc_chan_destroy(self._h); self._h = None

    @property
    def blockcount(self):
        return lib().fdc_chan_blockcount(self._h)

    def reset(self):
        check(lib().fdc_chan_reset(self._h))

    def seek(self, first_block, history=None):
        check(lib().fdc_chan_seek(self._h, int(first_block)))
        if history is not None and self.ovl:
            h = np.ascontiguousarray(history, dtype=np.complex64)
            assert h.size == self.ovl
            check(lib().fdc_chan_set_history(self._h, _ptr(h)))

    def work_host(self, x, want_spectrum=False, outs=None):
        """x: complex64 array of nblocks*hop new samples -> (list of per-channel arrays, spectrum or None)."""
        x = np.ascontiguousarray(x, dtype=np.complex64)
        nblocks = x.size // self.hop
        if outs is None:
            slab = np.empty(int(self.lout_prefix[-1]) * nblocks, dtype=np.complex64)
            outs = [slab[int(self.lout_prefix[i]) * nblocks: int(self.lout_prefix[i + 1]) * nblocks] for i in range(self.nchan)]
        ptrs = (C.c_void_p * max(self.nchan, 1))(*[o.ctypes.data for o in outs])
        spec = np.empty(nblocks * self.N, dtype=np.complex64) if want_spectrum else None
        check(lib().fdc_chan_work_host(self._h, _ptr(x), nblocks, C.cast(ptrs, C.c_void_p) if self.nchan else None,
                                       _ptr(spec) if want_spectrum else None), "fdc_chan_work_host")
        return outs, spec

    def work_device(self, d_in, nblocks, d_out, d_spectrum=0, stream=0):
        """Raw device pointers (ints), e.g. torch tensor .data_ptr(); only enqueues."""
        check(lib().fdc_chan_work_device(self._h, C.c_void_p(d_in), int(nblocks), C.c_void_p(d_out) if d_out else None,
                                         C.c_void_p(d_spectrum) if d_spectrum else None, C.c_void_p(stream) if stream else None),
              "fdc_chan_work_device")

    def work_device_slab(self, d_in, nblocks, d_out, slab_blocks, slab_first_block, d_spectrum=0, stream=0):
        """work_device with explicit output placement (rows [slab_first_block, +nblocks) of slabs of slab_blocks rows);
        d_out may be peer memory of the sink rank (FDC.sharded.PeerSink)."""
        check(lib().fdc_chan_work_device_slab(self._h, C.c_void_p(d_in), int(nblocks), C.c_void_p(d_out) if d_out else None, int(slab_blocks),
                                              int(slab_first_block), C.c_void_p(d_spectrum) if d_spectrum else None,
                                              C.c_void_p(stream) if stream else None), "fdc_chan_work_device_slab")

    def set_sinks(self, bases, owner, local_sink=-1):
        """channel-sharded sinks: bases[k] = device address of sink k's buffer as seen from this GPU, owner[i] = sink of channel i,
        local_sink = the sink that lies in this GPU's own memory"""
        arr = (C.c_void_p * len(bases))(*[int(b) if b else None for b in bases])
        own = (C.c_int * len(owner))(*[int(o) for o in owner])
        check(lib().fdc_chan_set_sinks(self._h, len(bases), C.cast(arr, C.c_void_p), C.cast(own, C.c_void_p), int(local_sink)), "fdc_chan_set_sinks")

    def work_device_sinks(self, d_in, nblocks, slab_blocks, slab_first_block, stream=0):
        """like work_device_slab, every channel's rows going to its owner's sink (FDC.sharded.ChannelSinks)"""
        check(lib().fdc_chan_work_device_sinks(self._h, C.c_void_p(d_in), int(nblocks), int(slab_blocks), int(slab_first_block),
                                               C.c_void_p(stream) if stream else None), "fdc_chan_work_device_sinks")

    def work_spectrum_device(self, d_spectra, nblocks, d_out, d_spectrum=0, stream=0):
        """inpveclen > 1 mode: already transformed (fft-shifted, unnormalised) spectra in, channel outputs out."""
        check(lib().fdc_chan_work_spectrum_device(self._h, C.c_void_p(d_spectra), int(nblocks), C.c_void_p(d_out) if d_out else None,
                                                  C.c_void_p(d_spectrum) if d_spectrum else None, C.c_void_p(stream) if stream else None),
              "fdc_chan_work_spectrum_device")

    def sync(self):
        check(lib().fdc_chan_sync(self._h))

    def set_profiling(self, enable):
        check(lib().fdc_chan_set_profiling(self._h, int(bool(enable))))

    def get_profile(self):
        """(ms in forward-FFT kernels, ms in extract kernels, chunks) since the last call; synchronises."""
        a, b, n = C.c_double(), C.c_double(), C.c_long()
        check(lib().fdc_chan_get_profile(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    @property
    def chunk_blocks(self):
        return lib().fdc_chan_chunk_blocks(self._h)

    @chunk_blocks.setter
    def chunk_blocks(self, v):
        check(lib().fdc_chan_set_chunk_blocks(self._h, int(v)))

    def out_slices(self, nblocks):
        """(offset, length) in items of every channel inside the device slab of a call with nblocks blocks."""
        return [(int(self.lout_prefix[i]) * nblocks, self.lout[i] * nblocks) for i in range(self.nchan)]
