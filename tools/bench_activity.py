#!/usr/bin/env python
"""measurement aid: the activity-gated blocks on a cfg3-style spectrum stream: FFT 16384, R = 4, bursty DAMA carriers
in two detection segments + 16 power-activated channels, spectra resident in device memory (the hier block's wiring).
Prints blocks/s and the equivalent input Msamples/s (hop = 12288 samples per block)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import FDC
import scenarios as sc

N, R = 16384, 4
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, truth = sc.bursty_spectra(N, nblocks, 48, seed=3, widths=(16, 32, 64, 128), raster=256, mean_on=24, mean_off=40)
L = FDC._cabi.lib()
d = L.fdc_dev_alloc(8 * x.size)
FDC._cabi.check(L.fdc_memcpy_h2d(d, x.ctypes.data, 8 * x.size))
segs = [FDC.SegmentDetection(i, N, R, a, b, 10.0, 0.002, 0.2, 128, 1, True, False, "", False, 0) for i, (a, b) in enumerate([(0.1, 0.45), (0.55, 0.9)])]
starts = sorted(set(t[0] for t in truth))[:16]
pacs = [FDC.PowerActivationChannel(N, (s + 32) / float(N), 64.0 / N, R, 6.0, 128, 1, True, False, "", 0, i) for i, s in enumerate(starts)]
for name, blocks_ in (("SegmentDetection x2", segs), ("PowerActivationChannel x16", pacs), ("all", segs + pacs)):
    for chunk in (64, 256, nblocks):
        nmsg = 0; nsamp = 0
        t0 = time.perf_counter()
        for b0 in range(0, nblocks, chunk):
            nb = min(chunk, nblocks - b0)
            for blk in blocks_:
                blk.work_device(nb, d + 8 * b0 * N)
                for m in blk.messages():
                    nmsg += 1; nsamp += m["nsamples"]
        L.fdc_device_synchronize()
        dt = time.perf_counter() - t0
        print("%-28s chunk %5d: %8.0f blocks/s = %8.1f Msamples/s  (%d PDUs, %.1f Msamples of bursts)" % (name, chunk, nblocks / dt, nblocks * (N - N // R) / dt / 1e6, nmsg, nsamp / 1e6))
