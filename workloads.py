"""Synthetic workloads of BASELINE.json's configs: geometry + seeded input generators.

Shared by tests/, bench.py and __graft_entry__.smoke().  Pure numpy; nothing here touches the GPU or the oracle.
Channel geometry comes from the hier block's get_opt_channelparams (python/FrequencyDomainChannelizer.py:322-345),
restated in geometry.py.
"""
import numpy as np

import geometry

RECTANGULAR, HANN, RAMP = 0, 1, 2


class ChanConfig(object):
    """A fixed-channel (throughput) configuration of the hier block."""

    def __init__(self, name, N, R, user_channels, windowtype, ovl=None):
        self.name, self.N, self.R, self.windowtype = name, int(N), int(R), int(windowtype)
        self.ovl = self.N // self.R if ovl is None else int(ovl)
        self.hop = self.N - self.ovl
        self.user_channels = list(user_channels)
        self.params = [geometry.get_opt_channelparams(self.N, self.R, geometry.get_freq(f), geometry.get_bw(bw))
                       for (f, bw) in self.user_channels]      # (f, l, lout, passbw, stopbw)
        if ovl is not None and self.ovl != self.N // self.R:
            # true sliding-window overlap (not expressible in the reference's hier block): keep l - l*ovl/N samples
            self.params = [(f, l, l - (l * self.ovl) // self.N, pb, sb) for (f, l, lo, pb, sb) in self.params]
        self.nchan = len(self.params)
        self.out_per_block = sum(p[2] for p in self.params)

    # shift of the phase table per block (phase_shifting_windowing_vcc `shifts` argument)
    def shifts(self):
        if self.ovl == self.N // self.R:
            return [p[0] for p in self.params]                                       # the hier block passes f
        # general hop: phase advance per block is -2 pi k0 hop / N; with nphase = R states that is shift = -f*hop*R/N
        return [(-p[0] * self.hop * self.R // self.N) for p in self.params]

    def bytes_per_sample(self):
        return 8.0 + 8.0 * self.out_per_block / self.hop

    def flops_per_sample(self):
        fl = 5.0 * self.N * np.log2(self.N) + sum(5.0 * p[1] * np.log2(p[1]) for p in self.params)
        return fl / self.hop


def cfg1():
    """FDC_example-style: FFT 1024, 50 % overlap, 16 fixed channels, multitone."""
    ch = [((i + 0.5) / 16.0 - 0.5, 0.05) for i in range(16)]
    return ChanConfig("cfg1_fft1024_r2_16ch", 1024, 2, ch, RECTANGULAR)


def cfg2(windowtype=HANN):
    """FFT 8192, 64 equal-bandwidth channels, shaped transition band."""
    ch = [((i + 0.5) / 64.0 - 0.5, 1.0 / 64.0) for i in range(64)]
    return ChanConfig("cfg2_fft8192_r4_64ch", 8192, 4, ch, windowtype)


def cfg4(true_overlap=False):
    """FFT 65536, 256 channels; R = 4 reading (25 % overlap, what the hier block can express) or a true 75 % overlap."""
    ch = [((i + 0.5) / 256.0 - 0.5, 1.0 / 256.0) for i in range(256)]
    if true_overlap:
        return ChanConfig("cfg4_fft65536_ovl75_256ch", 65536, 4, ch, HANN, ovl=49152)
    return ChanConfig("cfg4_fft65536_r4_256ch", 65536, 4, ch, HANN)


def cfg5_fixed():
    """FFT 262144, 4096 narrow channels (all active) -- throughput reading of config 5."""
    ch = [((i + 0.5) / 4096.0 - 0.5, 1.0 / 4096.0) for i in range(4096)]          # 64-bin raster: l = 128, lout = 96
    return ChanConfig("cfg5_fft262144_r4_4096ch", 262144, 4, ch, HANN)


def tones_input(cfg, nsamples, seed, noise=1e-2):
    """One tone per channel, slightly off centre, plus complex white noise (SURVEY 8d cfg1/cfg4 recipe)."""
    rng = np.random.default_rng(seed)
    n = np.arange(nsamples, dtype=np.float64)
    x = np.zeros(nsamples, dtype=np.complex128)
    nch = len(cfg.user_channels)
    # a sum over up to 4096 tones is built block-wise in the frequency domain to stay cheap
    for (fc, bw) in cfg.user_channels[:: max(1, nch // 64)]:
        d = rng.uniform(-0.2, 0.2) * bw
        a = rng.uniform(0.5, 1.0)
        x += a * np.exp(2j * np.pi * (fc + d) * n + 1j * rng.uniform(0, 2 * np.pi))
    x += noise * (rng.standard_normal(nsamples) + 1j * rng.standard_normal(nsamples))
    return x.astype(np.complex64)


def noise_input(nsamples, seed):
    """Complex white noise, unit variance per component (device-side bench input has the same statistics)."""
    rng = np.random.default_rng(seed)
    x = np.empty(nsamples, dtype=np.complex64)
    x.real = rng.standard_normal(nsamples, dtype=np.float32)
    x.imag = rng.standard_normal(nsamples, dtype=np.float32)
    return x


def example_channels():
    """examples/FDC_example.grc:143 -- channels with every phase shift 0..3."""
    return [(0.12, 0.05), (0.22, 0.1), (-0.14, 0.12), (0.0, 0.081)]


def cfg_example(N=4096, R=4, windowtype=RECTANGULAR):
    return ChanConfig("example_fft%d_r%d_4ch" % (N, R), N, R, example_channels(), windowtype)
