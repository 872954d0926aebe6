#!/usr/bin/env python
"""measurement aid: throughput of the fused channelizer vs the number of blocks per call (a GNU Radio scheduler hands a block
small batches unless set_output_multiple / set_min_output_buffer ask for more)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ctypes
import numpy as np
import torch
import FDC
import workloads
from helpers import make_gpu_chain

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cfg = {"cfg4": workloads.cfg4, "cfg2": workloads.cfg2, "cfg1": workloads.cfg1}[wl]()
L = FDC._cabi.lib()
maxb = 256
d_in = torch.randn(maxb * cfg.hop * 2, dtype=torch.float32, device="cuda")
d_out = torch.empty(maxb * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
h_in = L.fdc_host_alloc(8 * maxb * cfg.hop); h_out = L.fdc_host_alloc(8 * maxb * cfg.out_per_block)
chan = make_gpu_chain(FDC, cfg)
stream = torch.cuda.current_stream().cuda_stream
for nb in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    reps = max(4, 2048 // nb)
    for _ in range(3):
        chan.work_device(d_in.data_ptr(), nb, d_out.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        chan.work_device(d_in.data_ptr(), nb, d_out.data_ptr(), 0, stream)
    torch.cuda.synchronize()
    dt_dev = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):                                   # one call, result needed before the next (latency)
        chan.work_device(d_in.data_ptr(), nb, d_out.data_ptr(), 0, stream); torch.cuda.synchronize()
    dt_lat = (time.perf_counter() - t0) / reps
    outs = []; off = 0
    for lo in chan.lout:
        outs.append(h_out + off); off += 8 * nb * lo
    ptrs = (ctypes.c_void_p * len(outs))(*outs)
    reps_h = max(3, reps // 8)
    FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(h_in), nb, ctypes.cast(ptrs, ctypes.c_void_p), None))
    t0 = time.perf_counter()
    for _ in range(reps_h):
        FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(h_in), nb, ctypes.cast(ptrs, ctypes.c_void_p), None))
    dt_host = (time.perf_counter() - t0) / reps_h
    s = nb * cfg.hop / 1e6
    print("%s  %3d blocks/call: device back-to-back %8.0f Ms/s (%.0f us/call)   device call+sync %8.0f Ms/s (%.0f us)   host buffers %6.0f Ms/s (%.0f us)" %
          (wl, nb, s / dt_dev, dt_dev * 1e6, s / dt_lat, dt_lat * 1e6, s / dt_host, dt_host * 1e6))
