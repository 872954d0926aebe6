"""Test helpers: build the same fixed-channel configuration on the oracle and on the CUDA library."""
import numpy as np


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel(); b = np.asarray(b).astype(np.complex128).ravel()
    assert a.shape == b.shape
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def make_ref_chain(ref, cfg):
    """oracle: the hier block's throughput flowgraph out of the unmodified reference blocks"""
    assert cfg.ovl == cfg.N // cfg.R, "the reference's overlap_save only supports overlap N/R"
    return ref.Chain(cfg.N, cfg.R, cfg.params, cfg.windowtype)


def make_gpu_chain(FDC, cfg):
    chans = []
    for (f, l, lout, pb, sb), shift in zip(cfg.params, cfg.shifts()):
        table = FDC.psw_tables(l, cfg.R, pb, sb, cfg.windowtype)
        chans.append((f, l, lout, shift, float(l), table))
    return FDC.Channelizer(cfg.N, cfg.ovl, cfg.R, chans)


def run_chunked(work, x, hop, chunks):
    """feed x through work(samples) in the given block chunking, concatenating per-channel outputs"""
    outs = None; pos = 0
    for nb in chunks:
        o = work(x[pos * hop:(pos + nb) * hop]); pos += nb
        outs = [list(o)] if outs is None else outs + [list(o)]
    return [np.concatenate([c[i] for c in outs]) for i in range(len(outs[0]))]
