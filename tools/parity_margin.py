#!/usr/bin/env python
"""measurement aid: worst per-channel rel-L2 of the parity cases against the compiled reference (margin to the 1e-5 bar)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import FDC
import workloads
from oracle import fdc_ref
from helpers import rel_l2, make_ref_chain, make_gpu_chain
fdc_ref.set_fft_mode(0)
for name, mk, nb in (("cfg1", workloads.cfg1, 40), ("cfg2", workloads.cfg2, 6), ("cfg4", workloads.cfg4, 4),
                     ("example32768", lambda: workloads.cfg_example(32768, 4, workloads.HANN), 5)):
    cfg = mk()
    x = workloads.tones_input(cfg, nb * cfg.hop, seed=11) + workloads.noise_input(nb * cfg.hop, 12) * np.float32(0.05)
    want, wspec = make_ref_chain(fdc_ref, cfg).run(x, nthreads=8, want_spectrum=True)
    outs, spec = make_gpu_chain(FDC, cfg).work_host(x, want_spectrum=True)
    errs = [rel_l2(o, w) for o, w in zip(outs, want)]
    print("%-14s spectrum %.2e  channels worst %.2e median %.2e" % (name, rel_l2(spec, wspec), max(errs), float(np.median(errs))))
