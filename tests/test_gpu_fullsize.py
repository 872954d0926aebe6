"""GPU parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run 256-MiB batches in
seconds): chunking / call-splitting invariance (bit exact), linearity, tone placement (every channel's own tone comes out
with unit amplitude and a constant phase step, nothing leaks into the other channels), and agreement of a random subset of
(channel, block) cells with the fp64 restatement."""
import numpy as np
import pytest

import workloads
from helpers import rel_l2, make_gpu_chain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def FDC():
    import FDC as m
    return m


def _run_device(FDC, cfg, x, splits, chunk=None):
    import torch
    g = make_gpu_chain(FDC, cfg)
    if chunk:
        g.chunk_blocks = chunk
    d_in = torch.from_numpy(x.view(np.float32).copy()).cuda()
    outs = [[] for _ in range(cfg.nchan)]
    pos = 0
    for nb in splits:
        d_out = torch.empty(nb * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
        g.work_device(d_in.data_ptr() + 8 * pos * cfg.hop, nb, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        a = d_out.cpu().numpy().view(np.complex64)
        for i, (off, ln) in enumerate(g.out_slices(nb)):
            outs[i].append(a[off:off + ln].copy())
        pos += nb
    return [np.concatenate(o) for o in outs]


@pytest.mark.parametrize("mk,nblocks", [(workloads.cfg4, 300), (workloads.cfg2, 1200), (workloads.cfg1, 9000)])
def test_chunking_and_call_split_invariance(FDC, mk, nblocks):
    cfg = mk()
    x = workloads.noise_input(nblocks * cfg.hop, 31)
    a = _run_device(FDC, cfg, x, (nblocks,))
    b = _run_device(FDC, cfg, x, (1, 7, nblocks // 3, nblocks - 8 - nblocks // 3), chunk=37)
    for i in range(cfg.nchan):
        assert np.array_equal(a[i].view(np.uint32), b[i].view(np.uint32)), i


def test_cfg4_linearity_and_subset_against_fp64(FDC):
    from oracle import fdc_numpy as fnp
    cfg = workloads.cfg4()
    nblocks = 200
    x = workloads.noise_input(nblocks * cfg.hop, 41); y = workloads.noise_input(nblocks * cfg.hop, 42)
    ox = _run_device(FDC, cfg, x, (nblocks,)); oy = _run_device(FDC, cfg, y, (nblocks,))
    z = (np.float32(0.75) * x - np.float32(1.5) * y).astype(np.complex64)
    oz = _run_device(FDC, cfg, z, (nblocks,))
    for i in range(0, cfg.nchan, 17):
        assert rel_l2(oz[i], 0.75 * ox[i].astype(np.complex128) - 1.5 * oy[i].astype(np.complex128)) < 1e-5
    # a handful of channels over a window of blocks against the fp64 restatement (same arithmetic, exact)
    sel = [0, 1, 100, 255]
    b0, nb = 150, 6
    seg = x[(b0 * cfg.hop - cfg.ovl):(b0 + nb) * cfg.hop]
    c0 = [((b0 % cfg.R) * (((cfg.params[i][0] % cfg.R) + cfg.R) % cfg.R)) % cfg.R for i in sel]
    want, _ = fnp.channelize(seg[cfg.ovl:], cfg.N, cfg.R, [cfg.params[i] for i in sel], cfg.windowtype, hist=seg[:cfg.ovl].astype(np.complex128), counter0=c0)
    for k, i in enumerate(sel):
        lo = cfg.params[i][2]
        assert rel_l2(ox[i][b0 * lo:(b0 + nb) * lo], want[k]) < 1e-5


def test_cfg4_every_channel_passes_its_own_tone(FDC):
    """one tone per channel centre (SURVEY Appendix B.10 at full size): amplitude 1 after the settled first block, constant
    phase step across block seams, and the energy stays in the channel it belongs to"""
    cfg = workloads.cfg4()
    nblocks = 40
    n = np.arange(nblocks * cfg.hop, dtype=np.float64)
    probe = [3, 64, 128, 200, 254]
    x = np.zeros(n.size, dtype=np.complex128)
    for c in probe:
        fc = cfg.user_channels[c][0] + 26.0 / cfg.N          # bin centred (no leakage of the rectangular forward window)
        x += np.exp(2j * np.pi * fc * n)
    outs = _run_device(FDC, cfg, x.astype(np.complex64), (nblocks,))
    # bins of the tones in the fft-shifted spectrum; a channel may only carry energy if a tone falls inside its slice
    # [f, f + l) (channel 0's slice is clamped to the top of the band by get_opt_channelparams, so it sees channel 254's tone)
    tone_bins = [int(round((cfg.user_channels[c][0] + 0.5) * cfg.N)) % cfg.N + 26 for c in probe]
    for c in range(cfg.nchan):
        lo = cfg.params[c][2]
        y = outs[c][lo:]                                   # skip the first block (zero history)
        p = float(np.mean(np.abs(y) ** 2))
        if c in probe:
            assert abs(np.mean(np.abs(y)) - 1.0) < 2e-3, (c, np.mean(np.abs(y)))
            step = np.angle(y[1:] * np.conj(y[:-1]))
            assert np.std(step) < 2e-2, (c, np.std(step))  # phase continuous across all block seams
        elif not any(cfg.params[c][0] - 64 <= k < cfg.params[c][0] + cfg.params[c][1] + 64 for k in tone_bins):
            assert p < 1e-6, (c, p)                        # > 60 dB down when no tone is inside (or next to) the slice
