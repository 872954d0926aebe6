"""FDC.FrequencyDomainChannelizer -- host mirror of the reference's Python hier block
(python/FrequencyDomainChannelizer.py), same 25 positional constructor arguments as its GRC <make>
(grc/FDC_FrequencyDomainChannelizer.xml:7).

The reference wires stream_to_vector -> overlap_save -> fft_vcc -> multiply_const and then six blocks per throughput
channel, one PowerActivationChannel per activity controlled channel and one SegmentDetection per segment.  Here the
front end and all throughput channels are ONE fused GPU context (Channelizer, fdc_chan_*), and the activity-gated
blocks read the spectrum where it already lies in device memory.  Without the GNU Radio scheduler the block is driven by

    outs = blk.work(samples)      # samples: complex64, a whole number of hops (blocksize - blocksize/relinvovl)

which returns what the hier block's stream output ports would carry for those input items: one array per throughput
channel, preceded by the normalised spectrum vectors when `debug` is set (output port 0, :314-315).  PDUs of the
activity-gated blocks go to the handler set with set_msg_handler() and are also queued for messages().
"""
import numpy as np

from . import _cabi
from ._cabi import check, lib
from .blocks import Channelizer, opt_channelparams, psw_tables
from .activity import PowerActivationChannel, SegmentDetection


class FREQMODE:
    normalized, basebandfs, centerfreqfs = range(3)


class VERBOSEMODE:
    NOLOG, LOGTOCONSOLE, LOGTOFILE = range(3)


def nextpow2(k):
    if k < 1:
        raise ValueError('Cannot evaluate next power 2 of {}'.format(k))
    return 2 ** int(np.ceil(np.log2(k)))


def frequency_conversions(freqmode, fs, centerfrequency):
    """(mode, get_freq, set_freq, get_bw, set_bw) of the hier block: user frequencies / bandwidths to and from the block's own
    normalisation (0 .. 1 over the fft-shifted band), python/FrequencyDomainChannelizer.py:70-91; pinned by
    tests/golden/freqmodes.json (the reference's own lines, executed)."""
    get_freq = lambda f: (f + 0.5) % 1.0
    set_freq = lambda f: f - 0.5
    get_bw = lambda bw: bw % 1.0
    set_bw = lambda bw: bw
    if freqmode == FREQMODE.normalized or freqmode == 'normalized':
        mode = FREQMODE.normalized
    elif freqmode == FREQMODE.basebandfs or freqmode == 'basebandfs':
        mode = FREQMODE.basebandfs
        get_freq = lambda f: (f / fs + 0.5) % 1.0
        set_freq = lambda f: (f - 0.5) * fs
        get_bw = lambda bw: (bw / fs) % 1.0
        set_bw = lambda bw: bw * fs
    elif freqmode == FREQMODE.centerfreqfs or freqmode == 'centerfreqfs':
        mode = FREQMODE.centerfreqfs
        get_freq = lambda f: ((f - centerfrequency) / fs + 0.5) % 1.0
        set_freq = lambda f: (f - 0.5) * fs + centerfrequency
        get_bw = lambda bw: (bw / fs) % 1.0
        set_bw = lambda bw: bw * fs
    else:
        raise ValueError('Unknown Frequency mode. Exiting...')
    return mode, get_freq, set_freq, get_bw, set_bw


class FrequencyDomainChannelizer(object):
    def __init__(self, inptype, inpveclen, blocksize, relinvovl,
                 throughput_channels,
                 activity_controlled_channels,
                 act_contr_threshold,
                 fs, centerfrequency, freqmode,
                 windowtype,
                 msgoutput, fileoutput, outputpath,
                 threaded,
                 activity_detection_segments, act_det_threshold, minchandist,
                 act_det_deactivation_delay, minchanflankpuffer, verbose,
                 pow_act_deactivation_delay,
                 pow_act_maxblocks, act_det_maxblocks,
                 debug):
        self.verbose = int(verbose)
        self.itemsize = inptype
        self.debug = bool(debug)

        # frequency conversion lambdas, python/FrequencyDomainChannelizer.py:70-91
        self.freqmode, self.get_freq, self.set_freq, self.get_bw, self.set_bw = frequency_conversions(freqmode, fs, centerfrequency)

        self.throughput_channels = self._parse(throughput_channels, self.get_channel,
                                               'Throughput channels are invalid. Exiting...',
                                               'Cannot convert {} to channel. must be list or tuple with channel frequency and bandwidth. ')
        self.activity_controlled_channels = self._parse(activity_controlled_channels, self.get_channel,
                                                        'Activity controlled channels are invalid. Exiting...',
                                                        'Cannot convert {} to channel. must be list or tuple with channel frequency and bandwidth. ')
        self.activity_detection_segments = self._parse(activity_detection_segments, self.get_segment,
                                                       'Activity detection segments are invalid. Exiting...',
                                                       'Cannot convert {} to segment. must be list or tuple with channel start and stop frequency. ')

        self.inpveclen = int(inpveclen) if int(inpveclen) > 0 else 1
        self.blocksize = nextpow2(blocksize)
        self.relinvovl = nextpow2(relinvovl)
        self.ovllen = self.blocksize // self.relinvovl
        self.inpblocklen = self.blocksize - self.ovllen
        if self.inpveclen != 1 and self.inpveclen != self.blocksize:
            # the reference wires the input straight into multiply_const_cc(1/N, blocksize) (:284-290): any other vector
            # length makes GNU Radio refuse the connection
            raise ValueError('inpveclen must be 1 (sample stream) or blocksize (already transformed input vectors)')
        if self.itemsize != 8:
            raise ValueError('Unknown input type. ')           # the reference's float branch is dead code (:205-210)

        if self.verbose:
            self.log('\n' + '#' * 32 + '\n')
            self.log('# gr-FDC Frequency Domain Channelizer Runtime Information')
            self.log('\n' + '#' * 32 + '\n')
            self.log('Blocksize     = {}'.format(self.blocksize))
            self.log('InputVecLen   = {}'.format(self.inpveclen))
            self.log('Relinvovl     = {}'.format(self.relinvovl))
            self.log('Ovllen        = {}'.format(self.ovllen))
            self.log('MsgOutput     = {}'.format(msgoutput))
            self.log('FileOutput    = {}'.format(fileoutput))
            self.log('Outputpath    = {}'.format(outputpath))
            self.log('Threaded      = {}'.format(threaded))
            self.log('Debugoutput   = {}'.format(self.debug))
            self.log('\n' + '#' * 32 + '\n')
            self.log('# Throughput channels:         {}'.format(str(self.throughput_channels)))
            self.log('# Activity control channels:   {}'.format(str(self.activity_controlled_channels)))
            self.log('# Activity detection segments: {}'.format(str(self.activity_detection_segments)))
            self.log('\n' + '#' * 32 + '\n')

        # throughput channels: geometry + tables, :219-231
        self.channel_params = []
        chans = []
        for i, (freq, bw) in enumerate(self.throughput_channels):
            f, l, lout, pbw, sbw = self.get_opt_channelparams(freq, bw)
            dec = self.blocksize // l
            if self.verbose:
                self.log('# Throughput Channel {}: dec={}, f={}, l={}, lout={}, bw=({}, {})'.format(i, dec, f, l, lout, pbw, sbw))
            self.channel_params.append((f, l, lout, pbw, sbw))
            # multiply_const_cc(blocksize/dec) with Python-2 integer division == l
            chans.append((f, l, lout, f, float(self.blocksize // dec), psw_tables(l, self.relinvovl, pbw, sbw, int(windowtype))))
        self.N_throughput_channelizers = len(chans)
        self.front = Channelizer(self.blocksize, self.ovllen, self.relinvovl, chans)

        self._msgs = []
        self._handler = None
        self.msgoutput = bool(msgoutput)
        self.PowerActChans = []
        for i, (cfreq, bw) in enumerate(self.activity_controlled_channels):
            b = PowerActivationChannel(self.blocksize, cfreq, bw, self.relinvovl, float(act_contr_threshold), int(pow_act_maxblocks),
                                       int(pow_act_deactivation_delay) if int(pow_act_deactivation_delay) >= 0 else 0,
                                       bool(msgoutput), bool(fileoutput), str(outputpath), self.verbose, i)
            b.set_msg_handler(self._publish)
            self.PowerActChans.append(b)
        self.SegmentDetectionChans = []
        for i, (startf, stopf) in enumerate(self.activity_detection_segments):
            b = SegmentDetection(i, self.blocksize, self.relinvovl, startf, stopf, float(act_det_threshold), self.get_bw(minchandist),
                                 float(minchanflankpuffer) if 0.0 <= float(minchanflankpuffer) else 0.2, int(act_det_maxblocks),
                                 int(act_det_deactivation_delay) if int(act_det_deactivation_delay) >= 0 else 0,
                                 bool(msgoutput), bool(fileoutput), str(outputpath), bool(threaded), self.verbose)
            b.set_msg_handler(self._publish)
            self.SegmentDetectionChans.append(b)
        self._dbuf = {}

    # ---- helpers with the reference's names ----
    @staticmethod
    def _parse(lst, conv, err_type, err_item):
        out = []
        if lst is None:
            return out
        if not isinstance(lst, (list, tuple)):
            raise ValueError(err_type)
        for k in lst:
            c = conv(k)
            if c is None:
                raise ValueError(err_item.format(k))
            out.append(c)
        return out

    def get_opt_channelparams(self, freq, bw):
        return opt_channelparams(self.blocksize, self.relinvovl, freq, bw)

    def get_channel(self, c):
        if not isinstance(c, (list, tuple)) or len(c) != 2:
            return None
        return [self.get_freq(c[0]), self.get_bw(c[1])]

    def get_segment(self, c):
        if not isinstance(c, (list, tuple)) or len(c) != 2:
            return None
        return [self.get_freq(c[0]), self.get_freq(c[1])]

    def log(self, s):
        if self.verbose == VERBOSEMODE.LOGTOCONSOLE:
            print(str(s))
        elif self.verbose == VERBOSEMODE.LOGTOFILE:
            self.logtofile(s)

    def logtofile(self, s, end='\n'):
        if not hasattr(self, 'logfile'):
            self.logfile = 'gr-FDC.FreqDomChan.log'
            with open(self.logfile, 'w') as fh:
                fh.write('\n')
        with open(self.logfile, 'a') as fh:
            fh.write(str(s) + str(end))

    # ---- message plumbing (msg_connect to the hier block's "msgout" port) ----
    def set_msg_handler(self, fn):
        self._handler = fn

    def _publish(self, m):
        if not self.msgoutput:
            return
        if self._handler is not None:
            self._handler(m)
        else:
            self._msgs.append(m)

    def messages(self, clear=True):
        res = list(self._msgs)
        if clear:
            self._msgs = []
        return res

    # ---- streaming ----
    def _dev(self, key, nbytes):
        cur = self._dbuf.get(key)
        if cur is None or cur[1] < nbytes:
            if cur is not None:
                lib().fdc_dev_free(cur[0])
            p = lib().fdc_dev_alloc(nbytes)
            if not p:
                raise _cabi.FDCError(_cabi.last_error())
            self._dbuf[key] = (p, nbytes)
        return self._dbuf[key][0]

    def __del__(self):
        for p, _ in getattr(self, '_dbuf', {}).values():
            try:
                lib().fdc_dev_free(p)
            except Exception:
                pass

    def work(self, samples):
        x = np.ascontiguousarray(samples, dtype=np.complex64)
        per_item = self.inpblocklen if self.inpveclen == 1 else self.blocksize
        nblocks = x.size // per_item
        if nblocks * per_item != x.size:
            raise ValueError('input must be a whole number of blocks of {} samples'.format(per_item))
        need_spec = self.debug or self.PowerActChans or self.SegmentDetectionChans
        if not need_spec and self.inpveclen == 1:
            outs, _ = self.front.work_host(x)
            return outs
        N = self.blocksize
        d_in = self._dev('in', 8 * x.size)
        d_spec = self._dev('spec', 8 * nblocks * N)
        nout = int(self.front.lout_prefix[-1]) * nblocks
        d_out = self._dev('out', 8 * max(nout, 1))
        check(lib().fdc_memcpy_h2d(d_in, x.ctypes.data, 8 * x.size))
        if self.inpveclen == 1:
            self.front.work_device(d_in, nblocks, d_out if nout else 0, d_spec, 0)
        else:                                          # already transformed input vectors, :284-290
            self.front.work_spectrum_device(d_in, nblocks, d_out if nout else 0, d_spec, 0)
        self.front.sync()
        for b in self.PowerActChans:
            b.work_device(nblocks, d_spec)
        for b in self.SegmentDetectionChans:
            b.work_device(nblocks, d_spec)
        outs = []
        if self.debug:
            spec = np.empty(nblocks * N, dtype=np.complex64)
            check(lib().fdc_memcpy_d2h(spec.ctypes.data, d_spec, 8 * spec.size))
            outs.append(spec.reshape(nblocks, N))
        if nout:
            slab = np.empty(nout, dtype=np.complex64)
            check(lib().fdc_memcpy_d2h(slab.ctypes.data, d_out, 8 * nout))
            for off, ln in self.front.out_slices(nblocks):
                outs.append(slab[off:off + ln])
        return outs
