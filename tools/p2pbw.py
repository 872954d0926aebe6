#!/usr/bin/env python
"""measurement aid: peer-to-peer copy bandwidth between the GPUs of the box, one direction and both directions at once
(sizes the channel-sharded sinks: every GPU sends and receives at the same time).  One process, torch copies on one stream
per direction."""
import sys
import torch

n = torch.cuda.device_count()
size = 1 << 28                                     # 256 MiB
bufs = [torch.empty(size, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
dst = [[torch.empty(size, dtype=torch.uint8, device="cuda:%d" % i) for _ in range(n)] for i in range(n)]


def run(pairs, reps=10):
    """pairs: list of (src, dst) device indices copied concurrently"""
    streams = [torch.cuda.Stream(device="cuda:%d" % s) for s, _ in pairs]
    for k, (s, d) in enumerate(pairs):
        with torch.cuda.stream(streams[k]):
            dst[d][s].copy_(bufs[s], non_blocking=True)
    for i in range(n):
        torch.cuda.synchronize(i)
    ev = []
    for k, (s, d) in enumerate(pairs):
        with torch.cuda.stream(streams[k]):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(streams[k])
            for _ in range(reps):
                dst[d][s].copy_(bufs[s], non_blocking=True)
            e1.record(streams[k]); ev.append((e0, e1))
    for i in range(n):
        torch.cuda.synchronize(i)
    return [reps * size / (e0.elapsed_time(e1) * 1e-3) / 1e9 for e0, e1 in ev]


print("GPUs:", n)
if n >= 2:
    print("0 -> 1 alone              : %.0f GB/s" % run([(0, 1)])[0])
    r = run([(0, 1), (1, 0)])
    print("0 -> 1 and 1 -> 0 together: %.0f + %.0f GB/s" % (r[0], r[1]))
if n >= 4:
    pairs = [(s, d) for s in range(n) for d in range(n) if s != d]
    r = run(pairs, reps=4)
    per_src = [sum(v for v, (s, d) in zip(r, pairs) if s == i) for i in range(n)]
    print("all-to-all (%d GPUs): egress per GPU %s GB/s" % (n, ", ".join("%.0f" % v for v in per_src)))
