#!/usr/bin/env python
"""measurement aid: where a SegmentDetection work_device call spends its time (cfg3-style input)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import FDC
import scenarios as sc
N, R = 16384, 4
nblocks = 1024
x, truth = sc.bursty_spectra(N, nblocks, 48, seed=3, widths=(16, 32, 64, 128), raster=256, mean_on=24, mean_off=40)
L = FDC._cabi.lib()
d = L.fdc_dev_alloc(8 * x.size)
FDC._cabi.check(L.fdc_memcpy_h2d(d, x.ctypes.data, 8 * x.size))
b = FDC.SegmentDetection(0, N, R, 0.1, 0.45, 10.0, 0.002, 0.2, 128, 1, True, False, "", False, 0)
for rep in range(5):          # rep 0 is cold (pinned buffers, message arena and pending buffers grow to their working size)
    t0 = time.perf_counter()
    FDC._cabi.check(L.fdc_segdet_work_device(b._h, nblocks, d, None))
    t1 = time.perf_counter()
    n = L.fdc_segdet_msg_count(b._h)
    ms = b.messages()
    t2 = time.perf_counter()
    print("rep %d: work_device %.2f ms (%.1f us/block), %d msgs (%.1f MB) fetched in %.2f ms, launches so far %d" % (
        rep, (t1 - t0) * 1e3, (t1 - t0) / nblocks * 1e6, n, 8e-6 * sum(m["data"].size for m in ms), (t2 - t1) * 1e3, L.fdc_launch_count()))
