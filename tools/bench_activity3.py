#!/usr/bin/env python
"""measurement aid: wall-clock share of the parts of one single-GPU step of an activity workload of bench.py (cfg3 | cfg5_activity):
front end (overlap-save + forward FFT), work_device of the heaviest SegmentDetection, messages_arrays.
FDC_ACT_TIMING=1 adds the library's own phase split of work_device on stderr."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
import FDC

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_activity"
bench.ACT = bench.ACTIVITY[name]
nb = bench.ACT["blocks"]
N, R, hop, x, segs, pac = bench.cfg3_stream(nb)
L = FDC._cabi.lib()
front = FDC.Channelizer(N, N // R, R, [])
sd = FDC.SegmentDetection(*bench.sd_args(0, *segs[0]))
d_in = torch.from_numpy(x.view(np.float32).copy()).cuda()
d_spec = torch.empty(nb * N * 2, dtype=torch.float32, device="cuda")
for rep in range(6):
    t0 = time.perf_counter()
    front.work_device(d_in.data_ptr(), nb, 0, d_spec.data_ptr(), 0); front.sync()
    t1 = time.perf_counter()
    sd.work_device(nb, d_spec.data_ptr())
    t2 = time.perf_counter()
    recs, data, offsets = sd.messages_arrays(reuse=True)
    t3 = time.perf_counter()
    print("%s rep %d: front %.3f ms, work_device %.3f ms, messages_arrays %.3f ms (%d PDUs, %.1f MB)" % (
        name, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, recs.size, 8e-6 * data.size))
