"""Spectrum-domain test scenarios for the activity-gated blocks (inputs are the normalised, fft-shifted spectra the
hier block feeds them).  Appendix B.8/B.9 of SURVEY.md plus a seeded bursty-carrier generator (SURVEY 8d cfg3 style)."""
import numpy as np


def kat_spectra(N=256, nblocks=16, carriers=(), seed=7, sigma=0.01):
    """noise sigma per component + constant-amplitude carriers: (amp, bin_lo, bin_hi, blk_lo, blk_hi) inclusive blocks"""
    rng = np.random.default_rng(seed)
    x = (sigma * rng.standard_normal((nblocks, N)) + 1j * sigma * rng.standard_normal((nblocks, N))).astype(np.complex64)
    for (amp, lo, hi, b0, b1) in carriers:
        x[b0:b1 + 1, lo:hi] += np.complex64(amp)
    return x


def b8_input():
    return kat_spectra(256, 16, [(1.0, 105, 125, 4, 9)])


def b9_input():
    return kat_spectra(256, 20, [(1.0, 104, 128, 4, 9), (0.5, 160, 172, 6, 14)])


def bursty_spectra(N, nblocks, ncarriers, seed, raster=None, widths=(16, 32, 64), mean_on=10, mean_off=14, snr_db=25.0,
                   lo=0.1, hi=0.9):
    """DAMA-like carriers on a raster inside [lo, hi) of the band, geometric on/off bursts, random phases per bin/block."""
    rng = np.random.default_rng(seed)
    sigma = 0.01
    x = (sigma * rng.standard_normal((nblocks, N)) + 1j * sigma * rng.standard_normal((nblocks, N))).astype(np.complex64)
    raster = raster or max(widths) * 4
    slots = np.arange(int(lo * N) // raster + 1, int(hi * N) // raster - 1)
    rng.shuffle(slots)
    amp = sigma * np.sqrt(2.0) * 10 ** (snr_db / 20.0)
    truth = []
    for s in slots[:ncarriers]:
        w = int(rng.choice(widths)); start = int(s) * raster + (raster - w) // 2
        b = int(rng.integers(0, mean_off))
        while b < nblocks:
            on = 2 + int(rng.geometric(1.0 / mean_on)); off = 3 + int(rng.geometric(1.0 / mean_off))
            e = min(nblocks, b + on)
            ph = np.exp(2j * np.pi * rng.random((e - b, w))).astype(np.complex64)
            x[b:e, start:start + w] += np.complex64(amp) * ph
            truth.append((start, start + w, b, e - 1))
            b = e + off
    return x, truth


def group_power(x, start, D, M, mean=False):
    """decimated power exactly as the generic VOLK / scalar loops accumulate it: sequential fp32 additions"""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    seg = x[:, start:start + D * M].reshape(x.shape[0], M, D)
    re = seg.real.astype(np.float32); im = seg.imag.astype(np.float32)
    acc = np.zeros((x.shape[0], M), dtype=np.float32)
    for k in range(D):
        acc = (acc + (re[:, :, k] * re[:, :, k] + im[:, :, k] * im[:, :, k]).astype(np.float32)).astype(np.float32)
    if mean:
        acc = (acc * np.float32(1.0 / np.float32(D))).astype(np.float32)
    return acc


def band_power(x, m0, m1):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    re = x.real.astype(np.float32); im = x.imag.astype(np.float32)
    acc = np.zeros(x.shape[0], dtype=np.float32)
    for i in range(m0, m1):
        acc = (acc + (re[:, i] * re[:, i] + im[:, i] * im[:, i]).astype(np.float32)).astype(np.float32)
    return acc


def strip_time(msg_id):
    """IDs start with a wall-clock stamp %Y-%m-%d-%H-%M-%S (six dash separated fields); compare the rest"""
    return msg_id.split(".", 1)[1] if "." in msg_id else msg_id


def meta_tuple(m):
    """everything a PDU carries except the wall-clock prefix of the ID and the samples themselves"""
    return (strip_time(m["ID"]), bool(m["finalized"]), int(m["part"]), int(m["blockstart"]), int(m["blockend"]),
            int(m["vectorstart"]), int(m["vectorend"]), round(m["rel_bw"], 12), round(m["rel_cfreq"], 12),
            int(m.get("nsamples", m["data"].size)))
