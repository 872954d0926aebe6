"""Oracle (TEST INFRASTRUCTURE ONLY): fp64 NumPy restatement of gr-FDC's throughput path.

The reference blocks themselves are available to the tests as oracle/_ref/libfdc_ref.so (the unmodified
lib/*_impl.cc compiled against oracle/shim).  This module restates, independently and in double precision,
what that path computes, so that both fp32 implementations (the reference build with its FFT stand-in and the
CUDA kernels) can be bounded against one exact answer.  Parity pin: tests/golden/*.npz hold outputs of the
reference's own code run in this container (tests/golden/make_golden.py); tests/test_oracle_golden.py checks
this restatement against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import anything under oracle/.
The product (gr-fdc_b200/) never does.

Reference files followed (paths relative to /root/reference):
  lib/overlap_save_impl.cc:62-81          block b = [last ovl items | hop new items], zero history at start (:52)
  python/FrequencyDomainChannelizer.py:206 fft_vcc(N, forward, rectangular, shift=True): out = fftshift(FFT(x))
  python/FrequencyDomainChannelizer.py:216 multiply_const(1/N)
  lib/vector_cut_vxx_impl.cc:59-72        out = in[offset : offset + blocklen]
  lib/windows.h:41-124                    cr_win(): real mask in double, R phase-rotated copies cast to float
  lib/phase_shifting_windowing_vcc_impl.cc:41-86  table selection counter = (counter + shift) % R per block
  python/FrequencyDomainChannelizer.py:228 fft_vcc(l, inverse, rectangular, shift=True): IFFT(ifftshift-like half swap)
  python/FrequencyDomainChannelizer.py:221,231 multiply_const(blocksize/dec) with dec = blocksize/l  ->  * l
  python/FrequencyDomainChannelizer.py:322-345 get_opt_channelparams (restated in geometry.py, Python 2 semantics)
Third-party pieces restated here (not in /root/reference): GNU Radio 3.7.x gr::fft::fft_vcc (FFTW3f unnormalised
transforms, half swap on the forward output / the inverse input), blocks.multiply_const_cc.
"""
import numpy as np

RECTANGULAR, HANN, RAMP = 0, 1, 2


def overlap_save(x, outputlen, overlaplen, hist=None):
    """lib/overlap_save_impl.cc:62-81 -- returns (blocks[nblocks, outputlen], history for the next call)."""
    hop = outputlen - overlaplen
    nblocks = x.size // hop
    if hist is None:
        hist = np.zeros(overlaplen, dtype=x.dtype)                       # :52 zero-initialised history
    ext = np.concatenate([hist, x[:nblocks * hop]])
    idx = np.arange(nblocks)[:, None] * hop + np.arange(outputlen)[None, :]
    return ext[idx], ext[ext.size - overlaplen:] if overlaplen else ext[:0]


def vector_cut(v, offset, blocklen):
    """lib/vector_cut_vxx_impl.cc:67-68"""
    return v[..., offset:offset + blocklen]


def fft_vcc_forward_shift(blocks):
    """gr::fft::fft_vcc(N, True, ones, True): unnormalised forward FFT, output halves swapped (DC at N/2)."""
    return np.fft.fftshift(np.fft.fft(blocks, axis=-1), axes=-1)


def fft_vcc_inverse_shift(blocks):
    """gr::fft::fft_vcc(l, False, ones, True): halves of the INPUT swapped, then unnormalised backward FFT."""
    l = blocks.shape[-1]
    return np.fft.ifft(np.fft.ifftshift(blocks, axes=-1), axis=-1) * l


def window_mask(wintype, blocksize, passbw, stopbw):
    """lib/windows.h:41-124 with normalize = false: the real mask w_d (double), value 1/blocksize in the pass band."""
    passbw = float(np.float32(passbw)); stopbw = float(np.float32(stopbw))    # float arguments promoted to double (:50-51)
    if passbw >= 1.0:                                                         # phase_shifting_windowing_vcc_impl.cc:42-45
        passbw, stopbw, wintype = 1.0, 1.0, RECTANGULAR
    elif stopbw >= 1.0:
        stopbw = 1.0
    lowsamps = int((1.0 - stopbw) * blocksize) // 2                            # the cast binds before /2 (:50)
    highsamps = int(passbw * blocksize)
    rampsamps = (blocksize - 2 * lowsamps - highsamps) // 2
    v = 1.0 / blocksize
    w = np.full(blocksize, v, dtype=np.float64)
    if wintype in (HANN, RAMP):
        for i in range(lowsamps):
            w[i] = 0.0; w[blocksize - 1 - i] = 0.0
        for i in range(rampsamps):
            if wintype == RAMP:
                a = v * (i + 1) / (rampsamps + 1)                              # :103
            else:
                a = v * (-np.cos((i + 1) / (rampsamps + 1) * np.pi) / 2.0 + 0.5)   # :120-121
            w[lowsamps + i] = a; w[blocksize - lowsamps - 1 - i] = a
    else:
        for i in range(lowsamps + rampsamps // 2):                            # :86
            w[i] = 0.0; w[blocksize - 1 - i] = 0.0
    return w, lowsamps, highsamps, rampsamps


def psw_tables(blocksize, relinvovl, passbw, stopbw, wintype, exact=False):
    """R tables polar(w_d[k], 2 pi c_i / R), c_0 = 0, c_{i+1} = (c_i + 1) % R (lib/windows.h:62-78).
    exact=False rounds to complex64 like the reference's cast (:75); exact=True keeps double."""
    w, _, _, _ = window_mask(wintype, blocksize, passbw, stopbw)
    t = np.empty((relinvovl, blocksize), dtype=np.complex128)
    c = 0
    for i in range(relinvovl):
        phi = 2.0 * np.pi * c / relinvovl
        t[i].real = w * np.cos(phi); t[i].imag = w * np.sin(phi)               # std::polar(double, double)
        c = (c + 1) % relinvovl
    return t if exact else t.astype(np.complex64)


def channelize(x, N, R, params, wintype, hist=None, counter0=None, want_spectrum=False, ovl=None, shifts=None):
    """The hier block's chain for fixed channels (python/FrequencyDomainChannelizer.py:201-231, 284-315) in fp64.

    x: complex input stream; params: list of (f, l, lout, passbw, stopbw) from get_opt_channelparams.
    The table values are the reference's float32-rounded tables (they are inputs of the arithmetic, not results).
    ovl / shifts generalise the hier block (which only wires ovl = N/R and shifts = f): a true sliding window of any
    overlap with the per-block table advance given explicitly (used for the "true 75 % overlap" reading of config 4).
    Returns (list of per-channel output streams as complex128, spectrum[nblocks, N] or None)."""
    ovl = N // R if ovl is None else int(ovl)
    blocks, _ = overlap_save(np.asarray(x, dtype=np.complex128), N, ovl, hist)
    spec = fft_vcc_forward_shift(blocks) / N
    nblocks = spec.shape[0]
    outs = []
    for ci, (f, l, lout, pb, sb) in enumerate(params):
        tables = psw_tables(l, R, pb, sb, wintype).astype(np.complex128)
        sh = f if shifts is None else shifts[ci]
        shift = ((sh % R) + R) % R                                           # shifts argument = f (:226), made positive (:58)
        c0 = 0 if counter0 is None else counter0[ci]
        phase = (c0 + np.arange(nblocks) * shift) % R                         # counter = (counter + shift) % R (:82)
        seg = vector_cut(spec, f, l) * tables[phase]
        y = fft_vcc_inverse_shift(seg)
        outs.append((vector_cut(y, l - lout, lout) * float(l)).reshape(-1))
    return outs, (spec if want_spectrum else None)
