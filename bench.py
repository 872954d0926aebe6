#!/usr/bin/env python
"""bench.py -- input Msamples/s (cfp32) of the frequency-domain channelizer hot path and % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg1|cfg4_ovl75|cfg5] [--impl reference]

One "step" = one pass of the hot path (overlap-save staging -> forward FFT -> all channels: bin cut, filter/phase table,
inverse FFT, overlap discard) over one batch of `blocks_per_step` overlap-save blocks of synthetic input.
  value : whole-job input samples/s with the batch resident in HBM (device pointers through the C ABI)
  e2e   : the same through fdc_chan_work_host: pinned HOST input and output buffers, H2D/D2H inside the timed region
  roofline : dominant kernel, algorithmic bytes per launch / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline : the reference's own blocks (oracle/_ref, compiled from the unmodified sources) on the host cores
N > 1: one process per GPU (torchrun); the stream is time-sharded into contiguous runs of blocks, every rank recomputes
its own halo, no collective on the data path ("weak" scaling: per-GPU batch fixed).  The NCCL gather of the channel
outputs to rank 0 is measured separately and reported under "gather".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

import workloads  # noqa: E402

WORKLOADS = {
    "cfg1": workloads.cfg1, "cfg2": workloads.cfg2, "cfg4": workloads.cfg4,
    "cfg4_ovl75": lambda: workloads.cfg4(True), "cfg5": workloads.cfg5_fixed,
}
METRIC = "input Msamples/s (cfp32)"


_JSON_OUT = None


def claim_stdout():
    """ONE JSON line on stdout: libraries (NCCL's version banner, torch warnings) write to file descriptor 1 behind Python's
    back, so fd 1 is pointed at stderr for the run and the line goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n"); out.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, ~5 ms period)."""

    def __init__(self, index):
        threading.Thread.__init__(self); self.daemon = True
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz, self.ok = index, False, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit(); self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Best effort: run this rank on the CPUs NVML reports as local to its GPU, so that first-touch places the pinned host
    buffers on that NUMA node (with 8 ranks the host<->device copies otherwise cross the socket interconnect)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and w * 64 + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to gpu %d" % (len(cpus), index)
    except Exception as exc:
        return "not bound (%s)" % str(exc)[:80]
    return "not bound"


def blocks_per_step(cfg):
    """batch whose input alone (>= 256 MiB) exceeds the 126 MB L2, so consecutive steps cannot hit in cache"""
    return max(64, int(np.ceil((256 << 20) / (8.0 * cfg.hop))))


def cpu_reference_run(cfg, seconds_target=12.0):
    """The reference's CPU implementation of the path: the unmodified gr-FDC blocks (oracle/_ref) in the hier block's
    topology with the fp32 FFT stand-in, all host cores.  Bounded sample: one batch sized from a short calibration run
    (at most 1024 blocks, to bound host memory), repeated until about `seconds_target` seconds of CPU work are done."""
    from oracle import fdc_ref
    from helpers import make_ref_chain
    if cfg.ovl != cfg.N // cfg.R:
        return None
    cores = os.cpu_count() or 1
    fdc_ref.set_fft_mode(1)
    chain = make_ref_chain(fdc_ref, cfg)
    nb = max(cores, 8)
    x = workloads.noise_input(nb * cfg.hop, 99)
    t = time.perf_counter(); chain.run(x, nthreads=cores); dt = time.perf_counter() - t
    nb2 = int(min(max(nb, nb * 3.0 / max(dt, 1e-3)), 1024, (1 << 31) // (8 * cfg.N)))        # about 3 s per repetition
    x = workloads.noise_input(nb2 * cfg.hop, 98)
    reps = 0; total = 0.0
    while total < seconds_target and reps < 64:
        t = time.perf_counter(); chain.run(x, nthreads=cores); total += time.perf_counter() - t; reps += 1
    fdc_ref.set_fft_mode(0)
    return {"value": reps * nb2 * cfg.hop / total / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "sample": "%d x %d blocks (%d samples) of %s, unmodified gr-FDC blocks + fp32 FFT/VOLK stand-ins, %.1f s" %
                      (reps, nb2, reps * nb2 * cfg.hop, cfg.name, total)}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    from oracle import fdc_ref
    from helpers import make_ref_chain
    fdc_ref.set_fft_mode(1)
    chain = make_ref_chain(fdc_ref, cfg)
    nb = max(8 * cores, 128) if cfg.N * 8 * max(8 * cores, 128) < (1 << 30) else max(2 * cores, 16)     # bounded sample per step
    x = workloads.noise_input(nb * cfg.hop, 97)
    for _ in range(args.warmup):
        chain.run(x, nthreads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        chain.run(x, nthreads=cores)
    dt = time.perf_counter() - t
    v = args.steps * nb * cfg.hop / dt / 1e6
    sample = "%d blocks (%d samples) of %s per step, unmodified gr-FDC blocks (oracle/_ref) + fp32 FFT/VOLK stand-ins" % (nb, nb * cfg.hop, cfg.name)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg.name, "fft": cfg.N, "overlap": cfg.ovl, "channels": cfg.nchan, "blocks_per_step": nb},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# the activity-gated workloads: configs[2] (FFT 16384, 48 carriers, 2 segments, 16 watched channels) and the activity half of
# configs[4] (FFT 262144, ~3900 narrow DAMA carriers with 10 % duty on a 64-bin raster, one segment over the whole band)
ACTIVITY = {
    "cfg3": dict(name="cfg3_fft16384_r4_activity", N=16384, R=4, carriers=48, widths=(16, 32, 64, 128), raster=256, mean_on=24, mean_off=40,
                 lo=0.1, hi=0.9, segs=[(0.1, 0.45), (0.55, 0.9)], npac=16, minchandist=0.002, blocks=1024, seed=3),
    "cfg5_activity": dict(name="cfg5_fft262144_r4_activity", N=262144, R=4, carriers=3900, widths=(40,), raster=64, mean_on=4, mean_off=36,
                          lo=0.02, hi=0.98, segs=[(0.02, 0.98)], npac=0, minchandist=16.0 / 262144, blocks=64, seed=5),
}
ACT = ACTIVITY["cfg3"]


def sd_args(i, a, b):
    """SegmentDetection(ID, blocklen, relinvovl, start, stop, thresh dB, minchandist, flank puffer, maxblocks, delay, msg, file, path, threads, verbose)"""
    return (i, ACT["N"], ACT["R"], a, b, 10.0, ACT["minchandist"], 0.2, 128, 1, True, False, "", False, 0)


def cfg3_stream(nblocks, seed=None, rows=None):
    """Bursty DAMA carriers: the spectra are drawn per block and turned into a time stream block by block (the
    overlap-save blocks then see them smeared by the 25 % overlap, which is all a throughput measurement needs); the first
    `npac` carriers are also watched by PowerActivationChannels.  rows = (lo, hi): only the samples of blocks [lo, hi)."""
    import scenarios as sc
    N, R = ACT["N"], ACT["R"]
    hop = N - N // R
    spec, truth = sc.bursty_spectra(N, nblocks, ACT["carriers"], seed=ACT["seed"] if seed is None else seed, widths=ACT["widths"], raster=ACT["raster"],
                                    mean_on=ACT["mean_on"], mean_off=ACT["mean_off"], lo=ACT["lo"], hi=ACT["hi"])
    lo, hi = rows if rows is not None else (0, nblocks)
    t = np.fft.ifft(np.fft.ifftshift(spec[lo:hi], axes=1), axis=1).astype(np.complex64) * np.float32(N)
    x = np.ascontiguousarray(t[:, N - hop:]).reshape(-1)
    starts = sorted(set(tr[0] for tr in truth))[:ACT["npac"]]
    pac = [((s0 + 32) / float(N), 64.0 / N) for s0 in starts]
    return N, R, hop, x, list(ACT["segs"]), pac


def run_cfg3_sharded(args, rank, world, local):
    """configs[2] on N GPUs: the stream is time sharded, every rank runs overlap-save + forward FFT + the K3
    measurements on its own blocks, the compact detection records are all-gathered, the sequential bookkeeping of the 18 blocks
    is replicated, every rank extracts the bursts its blocks emitted and the sink rank publishes the PDUs
    (FDC/sharded.py: ShardedActivityGroup).  Weak scaling: nb blocks per GPU and step."""
    import torch
    import torch.distributed as dist
    import FDC
    from FDC import sharded
    from concurrent.futures import ThreadPoolExecutor
    nb = args.blocks or ACT["blocks"]
    W = max(args.warmup, 3); K = max(args.steps, 1)
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = FDC._cabi.lib(); FDC._cabi.check(L.fdc_set_device(local))
    total = nb * world
    first, count = sharded.partition(total, world)
    lo = max(first[rank] - 1, 0)                     # one block before the own run: a freshly activated channel takes it too
    N, R, hop, x, segs, pac = cfg3_stream(total, rows=(lo, first[rank] + count[rank]))
    nloc = first[rank] + count[rank] - lo
    front = FDC.Channelizer(N, N // R, R, [])
    sd = [FDC.SegmentDetection(*sd_args(i, a, b)) for i, (a, b) in enumerate(segs)]
    pc = [FDC.PowerActivationChannel(N, f, bw, R, 6.0, 128, 1, True, False, "", 0, i) for i, (f, bw) in enumerate(pac)]
    pool = ThreadPoolExecutor(max_workers=max(1, min(len(sd) + len(pc), (os.cpu_count() or 1) // world)))
    sink = None; owners = None
    mode = os.environ.get("FDC_BENCH_ACT_SINK", "owners")          # owners | rank0 | nccl
    if mode == "owners":
        try:      # block instances dealt out over the ranks: every rank receives and assembles the bursts of the instances it owns
            owners = sharded.PeerBuffers(256 << 20, rank, world)
        except Exception as exc:
            sys.stderr.write("peer buffers unavailable (%s), using the NCCL gather\n" % exc)
    elif mode == "rank0":
        try:      # the gather fused into the extract kernel: burst samples are stored straight into rank 0's buffer over NVLink
            sink = sharded.PeerSink(0, 0, rank, world, dst=0, nbytes=256 << 20)
        except Exception as exc:        # no peer access on this box: NCCL gather
            sys.stderr.write("peer sink unavailable (%s), using the NCCL gather\n" % exc)
    grp = sharded.ShardedActivityGroup(sd + pc, rank, world, dst=0, pool=pool, sink=sink, arrays=True, owners=owners)
    d_in = torch.from_numpy(x.view(np.float32).copy()).cuda(local)
    d_spec = torch.empty(nloc * N * 2, dtype=torch.float32, device=d_in.device)
    own = d_spec.data_ptr() + 8 * N * (first[rank] - lo)
    prev = d_spec.data_ptr() if first[rank] > lo else 0
    stats = {"pdus": 0, "samples": 0}

    def step():
        front.work_device(d_in.data_ptr(), nloc, 0, d_spec.data_ptr(), 0)
        front.sync()
        res = grp.work(total, own, prev)
        if res is not None:
            for r_ in res:
                if r_ is not None:
                    stats["pdus"] += int(r_[0].size); stats["samples"] += int(r_[1].size)

    for _ in range(W):
        step()
    torch.cuda.synchronize(); dist.barrier(); stats["pdus"] = stats["samples"] = 0
    grp.phase_seconds.clear()
    sampler = ClockSampler(local); sampler.start()
    l0 = L.fdc_launch_count()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    torch.cuda.synchronize(); dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=d_in.device)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    launches = torch.tensor([int(L.fdc_launch_count() - l0)], dtype=torch.int64, device=d_in.device)
    dist.all_reduce(launches)
    st = torch.tensor([stats["pdus"], stats["samples"]], dtype=torch.int64, device=d_in.device)
    dist.all_reduce(st)                                  # the PDUs are published on the ranks that own the block instances
    stats["pdus"], stats["samples"] = int(st[0].item()), int(st[1].item())
    sampler.stop_flag = True; sampler.join()
    if rank == 0:
        value = K * total * hop / dt / 1e6
        peaks, peak_src = measured_peaks()
        alg = 8.0 * total * hop + 8.0 * stats["samples"] / K
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": ACT["name"], "fft": N, "overlap": N // R, "hop": hop, "segments": segs,
                           "power_activation_channels": len(pac), "blocks_per_step_per_gpu": nb, "sharding": "time",
                           "pdus_per_step": stats["pdus"] / K, "burst_samples_per_step": stats["samples"] / K,
                           "sink_phase_ms_per_step": {k: round(v / K * 1e3, 3) for k, v in grp.phase_seconds.items()},
                           "exchange": "all-gather of the detection records (a few bytes per block); burst samples " +
                                       ("stored by the extract kernels straight into the buffer of the rank that owns the block instance (instance i -> rank i mod N, "
                                        "CUDA IPC peer memory), then a barrier; every rank assembles the PDUs of its instances" if owners is not None else
                                        "stored by the extract kernels straight into rank 0's buffer (CUDA IPC peer memory), then a barrier" if sink is not None
                                        else "gathered to rank 0 (NCCL)"),
                           "timing": "wall clock, max over ranks (host state machines are part of the path), barrier + device synchronise on both sides",
                           "l2_policy": "input %.0f MB per step and GPU" % (8e-6 * nb * hop)},
                "clocks": sampler.result(), "e2e": None, "gpu_launches": int(launches.item()),
                "roofline": {"bound": "hbm", "kernel": "forward_fft + host state machines", "achieved": K * alg / dt / 1e9,
                             "peak": peaks["hbm_gbs"] * world, "unit": "GB/s", "frac": K * alg / dt / 1e9 / (peaks["hbm_gbs"] * world), "traffic": None,
                             "peak_source": peak_src, "note": "host bound: the replicated bookkeeping does not shard, see DESIGN.md (activity-gated blocks)"},
                "cpu_baseline": None}
        emit(line)
    dist.destroy_process_group()
    return 0


def cfg3_cpu_reference(nb, N, R, hop, x, segs, pac, pool, seconds=10.0):
    """configs[2] on the host cores: the reference's overlap_save + restated fft_vcc on all cores, then the unmodified
    SegmentDetection x2 and PowerActivationChannel x16 blocks, one thread per block as under GNU Radio's scheduler."""
    from oracle import fdc_ref as ref
    ref.set_fft_mode(1)
    cores = os.cpu_count() or 1
    chain = ref.Chain(N, R, [], workloads.HANN)
    rsd = [ref.SegmentDetection(*sd_args(i, a, b)) for i, (a, b) in enumerate(segs)]
    rpc = [ref.PowerActivationChannel(N, f, bw, R, 6.0, 128, 1, True, False, "", 0, i) for i, (f, bw) in enumerate(pac)]
    t1 = time.perf_counter(); reps = 0
    while time.perf_counter() - t1 < seconds and reps < 256:
        _, sp = chain.run(x, nthreads=cores, want_spectrum=True, want_outputs=False)
        list(pool.map(lambda b: (b.work(sp), b.messages()), rsd + rpc))
        reps += 1
    dtc = time.perf_counter() - t1
    ref.set_fft_mode(0)
    return {"value": reps * nb * hop / dtc / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "reference",
            "sample": "%d x %d blocks, reference overlap_save + restated fft_vcc on all cores, then the unmodified SegmentDetection x%d and "
                      "PowerActivationChannel x%d blocks, one thread per block as under GNU Radio's thread-per-block scheduler, %.1f s" % (reps, nb, len(segs), len(pac), dtc)}


def run_cfg3_reference(args):
    """--impl reference --workload cfg3: the CPU arm alone (no GPU needed)"""
    from concurrent.futures import ThreadPoolExecutor
    nb = args.blocks or ACT["blocks"]
    N, R, hop, x, segs, pac = cfg3_stream(nb)
    pool = ThreadPoolExecutor(max_workers=min(len(segs) + len(pac), os.cpu_count() or 1))
    t0 = time.perf_counter()
    cpu = cfg3_cpu_reference(nb, N, R, hop, x, segs, pac, pool, seconds=max(3.0, 3.0 * args.steps))
    emit({"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": (time.perf_counter() - t0) * 1e3 / max(args.steps, 1), "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": ACT["name"], "fft": N, "overlap": N // R, "hop": hop, "segments": segs,
                     "power_activation_channels": len(pac), "blocks_per_step_per_gpu": nb},
          "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})
    return 0


def run_cfg3(args, rank, world, local):
    """configs[2]: the activity-gated path.  One step = nb blocks through overlap-save + forward FFT (spectrum stays in device
    memory) + 2 SegmentDetection + 16 PowerActivationChannel blocks.  The state machines of those blocks run on the host,
    so the step is timed on the wall clock with a device synchronise on both sides."""
    import torch
    import FDC
    nb = args.blocks or ACT["blocks"]
    W = max(args.warmup, 3); K = max(args.steps, 1)
    N, R, hop, x, segs, pac = cfg3_stream(nb)
    torch.cuda.set_device(local)
    L = FDC._cabi.lib(); FDC._cabi.check(L.fdc_set_device(local))
    front = FDC.Channelizer(N, N // R, R, [])
    sd = [FDC.SegmentDetection(*sd_args(i, a, b)) for i, (a, b) in enumerate(segs)]
    pc = [FDC.PowerActivationChannel(N, f, bw, R, 6.0, 128, 1, True, False, "", 0, i) for i, (f, bw) in enumerate(pac)]
    d_in = torch.from_numpy(x.view(np.float32).copy()).cuda(local)
    d_spec = torch.empty(nb * N * 2, dtype=torch.float32, device=d_in.device)
    stats = {"pdus": 0, "samples": 0}

    # GNU Radio runs every block in its own thread (thread-per-block scheduler): the 18 sink blocks work concurrently, each
    # single threaded.  Same here for both arms: one worker thread per block (the C ABI contexts are independent).
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=min(len(sd) + len(pc), os.cpu_count() or 1))

    def one(b):
        b.work_device(nb, d_spec.data_ptr())
        recs, data, offsets = b.messages_arrays(reuse=True)  # every PDU's metadata and samples, without a dict per PDU
        return recs.size, int(data.size)

    def step():
        front.work_device(d_in.data_ptr(), nb, 0, d_spec.data_ptr(), 0)
        front.sync()
        for n, smp in pool.map(one, sd + pc):
            stats["pdus"] += n; stats["samples"] += smp

    for _ in range(W):
        step()
    torch.cuda.synchronize(); stats["pdus"] = stats["samples"] = 0
    sampler = ClockSampler(local); sampler.start()
    l0 = L.fdc_launch_count()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = int(L.fdc_launch_count() - l0)
    sampler.stop_flag = True; sampler.join()
    value = K * nb * hop / dt / 1e6
    out_bytes = 8.0 * stats["samples"] / K
    peaks, peak_src = measured_peaks()
    alg = (8.0 * nb * hop + out_bytes)
    cpu = None if args.no_cpu else cfg3_cpu_reference(nb, N, R, hop, x, segs, pac, pool)
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ACT["name"], "fft": N, "overlap": N // R, "hop": hop, "segments": segs,
                       "power_activation_channels": len(pac), "blocks_per_step_per_gpu": nb,
                       "pdus_per_step": stats["pdus"] / K, "burst_samples_per_step": stats["samples"] / K,
                       "timing": "wall clock (host state machines are part of the path), device synchronised on both sides",
                       "l2_policy": "input %.0f MB per step" % (8e-6 * nb * hop)},
            "clocks": sampler.result(), "e2e": None, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "forward_fft + host state machines", "achieved": value * 1e6 * alg / (nb * hop) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": value * 1e6 * alg / (nb * hop) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "peak_source": peak_src, "note": "host bound: see DESIGN.md (activity-gated blocks)"},
            "cpu_baseline": cpu}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS) + sorted(ACTIVITY))
    ap.add_argument("--blocks", type=int, default=0, help="blocks per step per GPU (default: input >= 256 MiB)")
    ap.add_argument("--chunk", type=int, default=0, help="override blocks per K1->K2 round trip")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained leg")
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
    ap.add_argument("--no-sinks", action="store_true", help="N > 1: keep every rank's outputs local (compute-only weak scaling)")
    ap.add_argument("--no-readings", action="store_true", help="cfg4 only: skip the second reading of the config (true 75 %% overlap)")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and "FDC_COPY_THREADS" not in os.environ:
        # one process per GPU shares the host cores: size each process's staging copy pool for its share (read when the library loads)
        os.environ["FDC_COPY_THREADS"] = str(max(1, (os.cpu_count() or 1) // world - 1))
    if args.workload in ACTIVITY:
        global ACT
        ACT = ACTIVITY[args.workload]
        if args.impl == "reference":
            return run_cfg3_reference(args) if rank == 0 else 0
        if world > 1 or os.environ.get("FDC_BENCH_FORCE_SHARDED") == "1":       # the latter: the time-sharded call sequence on one rank (measurement aid, under torchrun)
            return run_cfg3_sharded(args, rank, world, local)
        return run_cfg3(args, rank, world, local) if rank == 0 else 0
    cfg = WORKLOADS[args.workload]()
    if args.impl == "reference":
        return run_reference(args, cfg, rank, world)
    W = max(args.warmup, 3); K = max(args.steps, 1)

    import torch
    import torch.distributed as dist
    import FDC
    from helpers import make_gpu_chain
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    FDC._cabi.check(FDC._cabi.lib().fdc_set_device(local))
    numa = bind_to_gpu_numa_node(local)        # pinned staging buffers of the e2e leg land on the GPU's own NUMA node
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"            # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = FDC._cabi.lib()

    nb = args.blocks or blocks_per_step(cfg)
    chan = make_gpu_chain(FDC, cfg)
    if args.chunk:
        chan.chunk_blocks = args.chunk
    # time sharding: rank r owns global blocks [r*K'*nb, ...) -- contiguous run, own halo (zeros here: synthetic stream)
    chan.seek(rank * (W + K) * nb)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    d_in = torch.randn(nb * cfg.hop * 2, dtype=torch.float32, device=dev, generator=gen)
    d_out = torch.empty(nb * cfg.out_per_block * 2, dtype=torch.float32, device=dev)
    # a dedicated (non-default) torch stream: the kernels are enqueued on it through the C ABI and the CUDA events
    # that time them are recorded on the same stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # N > 1: the outputs are part of the step.  Channel-sharded sinks (FDC.sharded.ChannelSinks): rank k owns 1/N of the channels,
    # every rank's extract kernel stores each channel's rows into its owner's buffer over NVLink peer memory (an all-to-all fused
    # into the kernel).  `value` is then the gather-inclusive rate; the compute-only rate is reported beside it.
    sinks = None; sinks_note = None
    if world > 1 and not args.no_sinks:
        from FDC import sharded
        try:
            sinks = sharded.ChannelSinks(chan, nb, rank, world)
        except Exception as exc:                          # no peer access on this box: fall back to local outputs, say so
            sinks_note = "channel-sharded sinks unavailable: " + str(exc)[:160]
        ok = torch.tensor([1 if sinks else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if sinks and int(ok.item()) == 0:
            sinks.close(); sinks = None; sinks_note = "channel-sharded sinks unavailable on another rank"

    def step_local():
        chan.work_device(d_in.data_ptr(), nb, d_out.data_ptr(), 0, stream)

    def step():
        if sinks:
            sinks.step(d_in.data_ptr(), nb, stream)
        else:
            step_local()

    def nvlink_bytes():
        """(tx, rx) data bytes of this GPU's NVLinks so far (nvidia-smi nvlink -gt d), None when unavailable"""
        try:
            import subprocess
            out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(local)], capture_output=True, text=True, timeout=20).stdout
            tx = rx = 0
            for ln in out.splitlines():
                ln = ln.strip()
                if "Data Tx" in ln:
                    tx += int(ln.split(":")[-1].split()[0])
                elif "Data Rx" in ln:
                    rx += int(ln.split(":")[-1].split()[0])
            return (tx * 1024, rx * 1024) if (tx or rx) else None
        except Exception:
            return None

    def bracket():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    bracket()
    sampler = ClockSampler(local); sampler.start()
    l0 = L.fdc_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    bracket()
    ms = e0.elapsed_time(e1)
    launches = int(L.fdc_launch_count() - l0)
    if len(sampler.samples) < 5:              # very short timed region: keep the same load running for the clock record
        t_end = time.time() + 0.5
        while time.time() < t_end:
            step(); torch.cuda.synchronize()
        clock_window = "timed region + 0.5 s of the same steps"
    else:
        clock_window = "timed region"
    sampler.stop_flag = True; sampler.join()
    clocks = sampler.result(); clocks["window"] = clock_window
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * K * nb * cfg.hop / (ms * 1e-3) / 1e6

    # ---- sustained leg: the same step repeated for >= 2 s of device time, clocks sampled inside it (the K-step region above is
    # a few milliseconds: a burst at boost clocks; this is the number to quote as sustained) ----
    sustained = None
    if not args.no_sustained:
        reps = max(K, int(np.ceil(args.sustain_seconds * 1e3 / max(ms / K, 1e-3))))
        bracket()
        nv0 = nvlink_bytes() if world > 1 else None
        ssamp = ClockSampler(local); ssamp.start()
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            step()
        s1.record()
        bracket()
        sms = s0.elapsed_time(s1)
        ssamp.stop_flag = True; ssamp.join()
        ts = torch.tensor([sms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sms = float(ts.item())
        sustained = {"value": world * reps * nb * cfg.hop / (sms * 1e-3) / 1e6, "unit": "Msamples/s", "steps": reps, "seconds": sms * 1e-3,
                     "clocks": ssamp.result()}
        nv1 = nvlink_bytes() if nv0 else None
        if nv0 and nv1:
            exp = 8.0 * reps * nb * cfg.out_per_block * (world - 1) / world if sinks else 0.0
            sustained["nvlink_rank0"] = {"tx_GBps": (nv1[0] - nv0[0]) / (sms * 1e-3) / 1e9, "rx_GBps": (nv1[1] - nv0[1]) / (sms * 1e-3) / 1e9,
                                         "expected_tx_GBps": exp / (sms * 1e-3) / 1e9, "source": "nvidia-smi nvlink -gt d, rank 0's GPU, over the sustained leg"}

    # ---- per-kernel roofline (separate pass with events around each kernel) ----
    peaks, peak_src = measured_peaks()
    chan.set_profiling(True)
    PK = max(3, min(K, 10))
    for _ in range(PK):
        step_local()
    ms_fwd, ms_ext, chunks = chan.get_profile()
    chan.set_profiling(False)
    in_bytes = 8.0 * nb * cfg.hop * PK
    out_bytes = 8.0 * nb * cfg.out_per_block * PK
    spec_bytes = 8.0 * nb * cfg.N * PK
    kern = {"forward_fft": {"ms": ms_fwd, "algorithmic_GB": in_bytes / 1e9, "GBps_algorithmic": in_bytes / ms_fwd / 1e6,
                            "GBps_incl_spectrum_write": (in_bytes + spec_bytes) / ms_fwd / 1e6},
            "channel_extract": {"ms": ms_ext, "algorithmic_GB": out_bytes / 1e9, "GBps_algorithmic": out_bytes / ms_ext / 1e6}}
    # forward_fft is ONE kernel for N <= 16384 and two (columns, rows) above; the dominant single kernel is the extract
    # unless one forward kernel alone takes longer
    fwd_kernels = 2 if cfg.N > 16384 else 1
    dom = "forward_fft" if ms_fwd / fwd_kernels >= ms_ext else "channel_extract"
    ach = kern[dom]["GBps_algorithmic"]
    kern[dom]["launches"] = int(chunks) * (fwd_kernels if dom == "forward_fft" else len(set(p[1] for p in cfg.params)))
    kern[dom]["avg_launch_ms"] = kern[dom]["ms"] / max(1, kern[dom]["launches"])
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one steady-state step captured IN the
    # pipeline (ncu --cache-control none --replay-mode application, all launches of a step; profiles/README.md), per launch
    traffic = None; traffic_step = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh).get(cfg.name, {})
        if tj.get(dom) and tj.get("blocks_per_step"):
            traffic_step = {k: tj[k]["dram_bytes_per_step"] * nb / tj["blocks_per_step"] for k in ("forward_fft", "channel_extract") if k in tj}
            traffic = traffic_step[dom] / max(1, kern[dom]["launches"] / PK)
    path_gbs = value * 1e6 / world * cfg.bytes_per_sample() / 1e9
    l2if = (8.0 * cfg.N * (4 if cfg.N > 16384 else 2) + 0.6 * 8.0 * sum(p[1] for p in cfg.params) + 8.0 * cfg.out_per_block) / cfg.hop
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": (out_bytes if dom == "channel_extract" else in_bytes) / max(1, kern[dom]["launches"]),
                "algorithmic_bytes_per_step": {"forward_fft": in_bytes / PK, "channel_extract": out_bytes / PK, "path": (in_bytes + out_bytes) / PK},
                "dram_bytes_per_step": traffic_step,
                "launches_per_step": launches / K, "kernels": kern,
                "path": {"bytes_per_sample": cfg.bytes_per_sample(), "achieved": path_gbs, "frac": path_gbs / peaks["hbm_gbs"],
                         "sustained_frac": (sustained["value"] * 1e6 / world * cfg.bytes_per_sample() / 1e9 / peaks["hbm_gbs"]) if sustained else None,
                         "frac_of_8TBps_nominal": path_gbs / 8000.0,
                         "flop_per_sample": cfg.flops_per_sample(), "tflops_fp32": value * 1e6 / world * cfg.flops_per_sample() / 1e12},
                # the resource that actually bounds the three-kernel structure: bytes the kernels move between the SMs and L2 per
                # input sample (samples in, four-step intermediate out and in for N >= 32768, spectrum out, slices in at the measured 60 % of
                # their nominal size thanks to L1 hits of overlapping slices, channel samples out) against the measured 29 B/clk/SM
                "sm_l2_interface": {"bytes_per_sample": l2if, "peak": 8200.0, "unit": "GB/s", "peak_source": "measured: tools/lsubench.cu, tools/l2bw.cu (profiles/)",
                                    "achieved": value * 1e6 / world * l2if / 1e9, "frac": value * 1e6 / world * l2if / 1e9 / 8200.0}}

    # ---- N > 1: the compute-only rate (outputs stay on the rank that made them) beside the gather-inclusive `value` ----
    compute_only = None
    if world > 1 and sinks:
        for _ in range(3):
            step_local()
        bracket()
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(K):
            step_local()
        c1.record(); bracket()
        tc = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        compute_only = {"value": world * K * nb * cfg.hop / (float(tc.item()) * 1e-3) / 1e6, "unit": "Msamples/s", "steps": K,
                        "note": "every rank keeps its outputs in its own HBM: no bytes cross NVLink"}

    # ---- e2e: host buffers through the C ABI ----
    e2e = None
    if not args.no_e2e:
        nb_e = nb                                                              # the same batch as the device-resident step
        nbytes_in = 8 * nb_e * cfg.hop; nbytes_out = 8 * nb_e * cfg.out_per_block
        import ctypes
        h_in = L.fdc_host_alloc(nbytes_in); h_out = L.fdc_host_alloc(nbytes_out)
        if not h_in or not h_out:
            raise SystemExit("pinned allocation failed: " + FDC._cabi.last_error())
        xin = np.ctypeslib.as_array(ctypes.cast(h_in, ctypes.POINTER(ctypes.c_float)), shape=(nb_e * cfg.hop * 2,))
        xin[:] = np.random.default_rng(5 + rank).standard_normal(xin.size, dtype=np.float32)
        outs = []
        off = 0
        for lo in chan.lout:
            outs.append(h_out + off); off += 8 * nb_e * lo
        ptrs = (ctypes.c_void_p * len(outs))(*outs)
        EK = max(3, min(K, 10))
        for _ in range(2):
            FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(h_in), nb_e, ctypes.cast(ptrs, ctypes.c_void_p), None))
        bracket()
        t0 = time.perf_counter()
        for _ in range(EK):
            FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(h_in), nb_e, ctypes.cast(ptrs, ctypes.c_void_p), None))
        bracket()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * EK * nb_e * cfg.hop / float(tt.item()) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": nbytes_in, "d2h_bytes_per_step": nbytes_out, "steps": EK, "blocks_per_step": nb_e,
               "api": "fdc_chan_work_host (pinned host in/out used in place, 4-slot H2D/compute/D2H pipeline)", "cpu_affinity": numa}
        # the same call with ordinary pageable buffers (what a GNU Radio scheduler hands to work()): the library stages them through
        # its own pinned slots with its copy pool
        x_pg = np.array(xin); o_pg = np.empty(nb_e * cfg.out_per_block * 2, dtype=np.float32)
        outs_pg = []; off = 0
        for lo in chan.lout:
            outs_pg.append(o_pg.ctypes.data + 4 * off); off += 2 * nb_e * lo
        ptrs_pg = (ctypes.c_void_p * len(outs_pg))(*outs_pg)
        for _ in range(2):
            FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(x_pg.ctypes.data), nb_e, ctypes.cast(ptrs_pg, ctypes.c_void_p), None))
        bracket()
        t0 = time.perf_counter()
        for _ in range(EK):
            FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(x_pg.ctypes.data), nb_e, ctypes.cast(ptrs_pg, ctypes.c_void_p), None))
        bracket()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        e2e["pageable"] = {"value": world * EK * nb_e * cfg.hop / float(tp.item()) / 1e6, "unit": "Msamples/s",
                           "api": "fdc_chan_work_host on pageable numpy buffers (staged through library-owned pinned slots, copy pool of %d threads + caller)" % L.fdc_copy_threads()}
        # ... and with those same buffers page-locked once by the caller (fdc_host_register): DMA in place, no staging
        reg = L.fdc_host_register(ctypes.c_void_p(x_pg.ctypes.data), x_pg.nbytes) == 0 and L.fdc_host_register(ctypes.c_void_p(o_pg.ctypes.data), o_pg.nbytes) == 0
        if reg:
            for _ in range(2):
                FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(x_pg.ctypes.data), nb_e, ctypes.cast(ptrs_pg, ctypes.c_void_p), None))
            bracket()
            t0 = time.perf_counter()
            for _ in range(EK):
                FDC._cabi.check(L.fdc_chan_work_host(chan._h, ctypes.c_void_p(x_pg.ctypes.data), nb_e, ctypes.cast(ptrs_pg, ctypes.c_void_p), None))
            bracket()
            tr = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tr, op=dist.ReduceOp.MAX)
            e2e["registered"] = {"value": world * EK * nb_e * cfg.hop / float(tr.item()) / 1e6, "unit": "Msamples/s",
                                 "api": "the same pageable buffers after fdc_host_register (page-locked in place once, as a flowgraph would do with its stream buffers)"}
            L.fdc_host_unregister(ctypes.c_void_p(x_pg.ctypes.data)); L.fdc_host_unregister(ctypes.c_void_p(o_pg.ctypes.data))
        del x_pg, o_pg
        L.fdc_host_free(h_in); L.fdc_host_free(h_out)

    # ---- optional: NCCL gather of the channel outputs to the sink rank ----
    gather = None
    if world > 1:
        bufs = [torch.empty_like(d_out) for _ in range(world)] if rank == 0 else None
        for _ in range(2):
            dist.gather(d_out, bufs, dst=0)
        bracket()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        GK = 5
        for _ in range(GK):
            step(); dist.gather(d_out, bufs, dst=0)
        g1.record(); bracket()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"value": world * GK * nb * cfg.hop / (float(tg.item()) * 1e-3) / 1e6, "unit": "Msamples/s",
                  "note": "compute + NCCL gather of every rank's output slab to rank 0 (sink ingest bound)"}

    # ---- gather fused into the extract kernel: every rank stores its outputs straight into the sink rank's buffer (peer
    # memory over NVLink, CUDA IPC), nothing is staged locally and no collective follows ----
    if world > 1:
        from FDC import sharded
        try:
            sink = sharded.PeerSink(cfg.out_per_block, nb, rank, world, dst=0)

            def fstep():
                chan.work_device_slab(d_in.data_ptr(), nb, sink.ptr, sink.slab_blocks, sink.first_block(), 0, stream)

            for _ in range(3):
                fstep()
            bracket()
            f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
            FK = max(5, min(K, 20))
            f0.record()
            for _ in range(FK):
                fstep()
            f1.record(); bracket()
            tf = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            gather["fused_peer_store"] = {"value": world * FK * nb * cfg.hop / (float(tf.item()) * 1e-3) / 1e6, "unit": "Msamples/s",
                                          "note": "extract kernels of all ranks store into rank 0's buffer (fdc_chan_work_device_slab on IPC-mapped "
                                                  "peer memory): compute and transfer overlap, no NCCL call on the path"}
            bracket()
            sink.close()
        except Exception as exc:                          # peer access not available on this box: keep the NCCL number
            gather["fused_peer_store"] = {"unavailable": str(exc)[:200]}

    # ---- the config's other reading (BASELINE configs[3] says "75 % overlap"; the reference can only express overlap 1/R):
    # a short leg of the true-75 %-overlap workload (hop = N/4: 3x the transforms per input sample, compute bound) in the same line
    readings = None
    if args.workload == "cfg4" and not args.no_readings:
        cfg_b = WORKLOADS["cfg4_ovl75"]()
        nb_b = blocks_per_step(cfg_b)
        chan_b = make_gpu_chain(FDC, cfg_b)
        d_in_b = torch.randn(nb_b * cfg_b.hop * 2, dtype=torch.float32, device=dev, generator=gen)
        d_out_b = torch.empty(nb_b * cfg_b.out_per_block * 2, dtype=torch.float32, device=dev)
        for _ in range(3):
            chan_b.work_device(d_in_b.data_ptr(), nb_b, d_out_b.data_ptr(), 0, stream)
        bracket()
        KB = max(5, min(K, 10))
        b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(KB):
            chan_b.work_device(d_in_b.data_ptr(), nb_b, d_out_b.data_ptr(), 0, stream)
        b1.record(); bracket()
        tb = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        vb = world * KB * nb_b * cfg_b.hop / (float(tb.item()) * 1e-3) / 1e6
        readings = {"R4_hop_49152 (value; what the reference's hier block can express)": value,
                    "true_75pct_overlap_hop_16384": {"value": vb, "unit": "Msamples/s", "steps": KB, "blocks_per_step_per_gpu": nb_b,
                                                     "flop_per_sample": cfg_b.flops_per_sample(),
                                                     "tflops_fp32": vb * 1e6 / world * cfg_b.flops_per_sample() / 1e12,
                                                     "hbm_frac": vb * 1e6 / world * cfg_b.bytes_per_sample() / 1e9 / peaks["hbm_gbs"]}}
        del chan_b, d_in_b, d_out_b

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_run(cfg)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": cfg.name, "fft": cfg.N, "overlap": cfg.ovl, "hop": cfg.hop, "channels": cfg.nchan,
                           "slice_len": cfg.params[0][1], "out_per_block": cfg.out_per_block, "blocks_per_step_per_gpu": nb,
                           "chunk_blocks": chan.chunk_blocks, "sharding": "time (contiguous runs of blocks per rank, halo recomputed)",
                           "outputs": ("channel-sharded sinks: rank k owns channels with owner k, all ranks' extract kernels store into the owners' "
                                       "buffers over NVLink peer memory inside the timed region (%d B of %d B per sample leave each GPU)" %
                                       (int(8 * cfg.out_per_block / cfg.hop * (world - 1) / world), int(8 * cfg.out_per_block / cfg.hop))) if sinks
                                      else ("local to each rank" + ("; " + sinks_note if sinks_note else "")),
                           "l2_policy": "inputs larger than L2 (%.0f MB in, %.0f MB out per step)" %
                                        (8e-6 * nb * cfg.hop, 8e-6 * nb * cfg.out_per_block)},
                "clocks": clocks, "sustained": sustained, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
        if readings:
            line["readings"] = readings
        if compute_only:
            line["compute_only"] = compute_only
        if gather:
            line["gather"] = gather
        emit(line)
    if sinks:
        bracket(); sinks.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
