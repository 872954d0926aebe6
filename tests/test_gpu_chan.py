"""GPU parity: the CUDA throughput channelizer and block replacements against the oracle (the unmodified
reference blocks + restated third-party stages), through the C ABI.  Tolerance: channel samples and spectra
rel-L2 <= 1e-5 (BASELINE.json north_star); byte-copy blocks and tables bit exact."""
import numpy as np
import pytest

import scenarios as sc
import workloads
from helpers import rel_l2, make_ref_chain, make_gpu_chain

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def FDC():
    import FDC as m
    return m


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("forward,shift", [(True, True), (False, True), (True, False), (False, False)])
def test_fft_vcc_stage(FDC, ref, n, forward, shift):
    rng = np.random.default_rng(n + 7 * forward + 3 * shift)
    nvec = max(3, 9000 // n)
    x = (rng.standard_normal(nvec * n) + 1j * rng.standard_normal(nvec * n)).astype(np.complex64)
    want = ref.fft_vcc(x, n, forward, shift)
    got = FDC.fft_vcc(n, forward, None, shift, 1).process(x)
    assert rel_l2(got, want) < 2e-6


def test_overlap_save_bit_exact(FDC, ref):
    # Appendix B.4 plus cfp32 geometry, chunked calls
    x = np.arange(1, 61, dtype=np.float32)
    a = ref.overlap_save(4, 8, 2); b = FDC.overlap_save(4, 8, 2)
    want = np.concatenate([a.work(x[:30]).view(np.float32), a.work(x[30:]).view(np.float32)])
    got = np.concatenate([b.process(x[:6]), b.process(x[6:18]), b.process(x[18:])])
    assert np.array_equal(got, want)
    assert list(got[:16]) == [0, 0, 1, 2, 3, 4, 5, 6, 5, 6, 7, 8, 9, 10, 11, 12]
    rng = np.random.default_rng(5)
    y = (rng.standard_normal(3072 * 7) + 1j * rng.standard_normal(3072 * 7)).astype(np.complex64)
    a = ref.overlap_save(8, 4096, 1024); b = FDC.overlap_save(8, 4096, 1024)
    want = a.work(y).view(np.complex64)
    got = np.concatenate([b.process(y[:3072 * 2]), b.process(y[3072 * 2:])])
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def test_vector_cut_bit_exact(FDC, ref):
    x = np.array([0, 1, 2, 3, 10, 11, 12, 13], dtype=np.float32)
    assert list(FDC.vector_cut_vxx(4, 4, 1, 2).process(x)) == [1, 2, 11, 12]          # Appendix B.5
    rng = np.random.default_rng(6)
    y = (rng.standard_normal(4096 * 5) + 1j * rng.standard_normal(4096 * 5)).astype(np.complex64)
    want = ref.vector_cut_vxx(8, 4096, 963, 1024).work(y).view(np.complex64)
    got = FDC.vector_cut_vxx(8, 4096, 963, 1024).process(y)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    with pytest.raises(FDC.FDCError):
        FDC.vector_cut_vxx(8, 16, 10, 8)


@pytest.mark.parametrize("args", [(200, 4, 3, 0.5, 0.75, 2), (256, 4, 5, 0.55, 0.8, 1), (64, 2, 7, 0.88, 1.0, 0),
                                  (512, 8, -3, 0.3, 0.9, 1)])
def test_psw_block_bit_exact(FDC, ref, args):
    a = ref.phase_shifting_windowing_vcc(*args); b = FDC.phase_shifting_windowing_vcc(*args)
    assert np.array_equal(a.tables().view(np.uint8), b.tables().view(np.uint8))
    rng = np.random.default_rng(args[0])
    x = (rng.standard_normal(args[0] * 11) + 1j * rng.standard_normal(args[0] * 11)).astype(np.complex64)
    want = np.concatenate([a.work(x[:args[0] * 2]).view(np.complex64), a.work(x[args[0] * 2:]).view(np.complex64)])
    got = np.concatenate([b.process(x[:args[0] * 5]), b.process(x[args[0] * 5:])])
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    assert a.state() == b.state()


CASES = {
    "cfg1": (workloads.cfg1, 40),
    "example4096_rect": (lambda: workloads.cfg_example(4096, 4, workloads.RECTANGULAR), 12),
    "example4096_hann": (lambda: workloads.cfg_example(4096, 4, workloads.HANN), 12),
    "example2048_r2_ramp": (lambda: workloads.cfg_example(2048, 2, workloads.RAMP), 12),
    "example16384_r8": (lambda: workloads.cfg_example(16384, 8, workloads.HANN), 6),
    "cfg2": (workloads.cfg2, 6),
    "example32768": (lambda: workloads.cfg_example(32768, 4, workloads.HANN), 5),
    "cfg4": (workloads.cfg4, 4),
    # BASELINE configs[4], throughput reading: FFT 262144 (512 x 512 four-step), 4096 channels of 128 bins in one launch
    "cfg5_fixed": (workloads.cfg5_fixed, 3),
    # narrow channels on long transforms (slices 128 .. 2048 bins): 256 x 512 and 512 x 512 four-step, 32 points per thread
    "narrow131072": (lambda: workloads.ChanConfig("narrow131072", 131072, 4, NARROW, workloads.HANN), 3),
    "narrow262144_r8": (lambda: workloads.ChanConfig("narrow262144", 262144, 8, NARROW, workloads.RAMP), 3),
    # the largest transforms the forward kernels take: 512 x 1024 and 1024 x 1024 four-step
    "narrow524288": (lambda: workloads.ChanConfig("narrow524288", 524288, 4, NARROW, workloads.HANN), 3),
    "narrow1048576_r2": (lambda: workloads.ChanConfig("narrow1048576", 1048576, 2, NARROW, workloads.RECTANGULAR), 3),
}
NARROW = [(0.12, 0.001), (0.22, 0.004), (-0.14, 0.0005), (0.0, 0.002), (-0.4991, 0.0003), (0.4993, 0.0003)]


@pytest.mark.parametrize("case", sorted(CASES))
def test_chain_matches_reference(FDC, ref, case):
    mk, nblocks = CASES[case]
    cfg = mk()
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=11) + workloads.noise_input(nblocks * cfg.hop, 12) * np.float32(0.05)
    if cfg.nchan > 1024:
        # tones_input puts a tone into every 64th channel only; the other 4032 channels of cfg5 would carry nothing but the 26 dB
        # weaker noise, and a per-channel RELATIVE error there measures the fp32 rounding noise the strong tones spread over all
        # 262144 bins (2e-5 of such a channel, for any fp32 transform) instead of the channel's own arithmetic.  "All channels
        # active" (BASELINE configs[4]): full-scale noise in every channel plus the tones.
        x = (workloads.noise_input(nblocks * cfg.hop, 12) + np.float32(0.25) * workloads.tones_input(cfg, nblocks * cfg.hop, seed=11)).astype(np.complex64)
    want, wspec = make_ref_chain(ref, cfg).run(x, nthreads=8, want_spectrum=True)
    g = make_gpu_chain(FDC, cfg)
    # two calls with different sizes: history and phase counters must carry over
    n1 = nblocks // 3
    o1, s1 = g.work_host(x[:n1 * cfg.hop], want_spectrum=True)
    o2, s2 = g.work_host(x[n1 * cfg.hop:], want_spectrum=True)
    spec = np.concatenate([s1, s2])
    assert rel_l2(spec, wspec) < TOL
    worst = 0.0
    for i in range(cfg.nchan):
        got = np.concatenate([o1[i], o2[i]])
        assert got.size == want[i].size == nblocks * cfg.params[i][2]
        worst = max(worst, rel_l2(got, want[i]))
    assert worst < TOL, worst
    assert g.blockcount == nblocks


@pytest.mark.parametrize("case,nblocks,chunk,splits", [
    ("cfg4", 9, 2, (9,)),                 # 1-block head + 4 chunks of 2 over the 3 worker streams, rings reused
    ("cfg4", 10, 3, (4, 6)),              # two calls: history and phase carried on the device
    ("cfg2", 25, 4, (25,)),               # single-kernel forward transform, 7 chunks
    ("cfg4_ovl50", 7, 1, (7,)),           # chunk of one block with a one-block head
])
def test_multichunk_device_path_matches_reference(FDC, ref, case, nblocks, chunk, splits):
    """the path bench.py times (fdc_chan_work_device: chunks alternating over the worker streams, per-stream mid / spectrum
    rings reused from chunk to chunk) against the compiled reference blocks, not only against itself"""
    import torch
    cfg = workloads.ChanConfig("cfg4_ovl50", 65536, 2, workloads.cfg4().user_channels[::16], workloads.RAMP) if case == "cfg4_ovl50" \
        else getattr(workloads, case)()
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=17) + workloads.noise_input(nblocks * cfg.hop, 18) * np.float32(0.05)
    want, _ = make_ref_chain(ref, cfg).run(x, nthreads=8)
    g = make_gpu_chain(FDC, cfg)
    g.chunk_blocks = chunk
    d_in = torch.from_numpy(x.view(np.float32).copy()).cuda()
    got = [[] for _ in range(cfg.nchan)]
    pos = 0
    for nb in splits:
        d_out = torch.empty(nb * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
        g.work_device(d_in.data_ptr() + 8 * pos * cfg.hop, nb, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        a = d_out.cpu().numpy().view(np.complex64)
        for i, (off, ln) in enumerate(g.out_slices(nb)):
            got[i].append(a[off:off + ln].copy())
        pos += nb
    worst = max(rel_l2(np.concatenate(got[i]), want[i]) for i in range(cfg.nchan))
    assert worst < TOL, worst


def test_host_buffers_pageable_pinned_registered_agree(FDC):
    """fdc_chan_work_host with the caller's buffers (a) pageable: staged through the library's pinned slots by the copy pool,
    (b) from fdc_host_alloc: used in place, (c) pageable but registered with fdc_host_register: used in place -- bit-identical
    outputs, several calls (history carried), more chunks than slots"""
    import ctypes
    cfg = workloads.cfg2()
    L = FDC._cabi.lib()
    calls = (3, 150, 40)                              # 150 blocks = 7 host chunks of 21 blocks on 4 slots
    x = workloads.noise_input(sum(calls) * cfg.hop, 77)

    def run(kind):
        g = make_gpu_chain(FDC, cfg)
        res = [[] for _ in range(cfg.nchan)]
        pos = 0
        for nb in calls:
            nin, nout = 8 * nb * cfg.hop, 8 * nb * cfg.out_per_block
            keep = []
            if kind == "pinned":
                h_in, h_out = L.fdc_host_alloc(nin), L.fdc_host_alloc(nout)
                assert h_in and h_out
            else:
                a = np.empty(nin, dtype=np.uint8); b = np.empty(nout, dtype=np.uint8); keep = [a, b]
                h_in, h_out = a.ctypes.data, b.ctypes.data
                if kind == "registered":
                    FDC._cabi.check(L.fdc_host_register(ctypes.c_void_p(h_in), nin)); FDC._cabi.check(L.fdc_host_register(ctypes.c_void_p(h_out), nout))
            ctypes.memmove(h_in, x[pos * cfg.hop:(pos + nb) * cfg.hop].ctypes.data, nin)
            outs = []; off = 0
            for lo in g.lout:
                outs.append(h_out + off); off += 8 * nb * lo
            ptrs = (ctypes.c_void_p * len(outs))(*outs)
            FDC._cabi.check(L.fdc_chan_work_host(g._h, ctypes.c_void_p(h_in), nb, ctypes.cast(ptrs, ctypes.c_void_p), None))
            got = np.ctypeslib.as_array(ctypes.cast(h_out, ctypes.POINTER(ctypes.c_uint8)), shape=(nout,)).copy().view(np.complex64)
            off = 0
            for i, lo in enumerate(g.lout):
                res[i].append(got[off:off + nb * lo]); off += nb * lo
            if kind == "pinned":
                L.fdc_host_free(h_in); L.fdc_host_free(h_out)
            elif kind == "registered":
                FDC._cabi.check(L.fdc_host_unregister(ctypes.c_void_p(h_in))); FDC._cabi.check(L.fdc_host_unregister(ctypes.c_void_p(h_out)))
            pos += nb
            del keep
        return [np.concatenate(r) for r in res]

    a, b, c = run("pageable"), run("pinned"), run("registered")
    for i in range(cfg.nchan):
        assert np.array_equal(a[i].view(np.uint32), b[i].view(np.uint32)), i
        assert np.array_equal(a[i].view(np.uint32), c[i].view(np.uint32)), i
    assert L.fdc_copy_threads() >= 0


def test_device_path_equals_host_path(FDC):
    import torch
    cfg = workloads.cfg2()
    nblocks = 700                                   # more than one L2 chunk (512 blocks)
    x = workloads.noise_input(nblocks * cfg.hop, 3)
    g1 = make_gpu_chain(FDC, cfg); g2 = make_gpu_chain(FDC, cfg)
    outs, _ = g1.work_host(x)
    d_in = torch.from_numpy(x.view(np.float32)).cuda()
    d_out = torch.empty(nblocks * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
    g2.work_device(d_in.data_ptr(), nblocks, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.complex64)
    for i, (off, ln) in enumerate(g2.out_slices(nblocks)):
        assert np.array_equal(got[off:off + ln].view(np.uint8), outs[i].view(np.uint8))


def test_golden_vectors_from_the_reference_build(FDC):
    """committed outputs of the reference's own code (tests/golden, made by make_golden.py) -- no oracle library needed"""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("chain_n1024_r4_hann", "chain_n512_r2_rect", "chain_n2048_r8_ramp"):
        g = np.load(os.path.join(gold, name + ".npz"))
        N, R, wintype, nblocks, seed = [int(v) for v in g["meta"]]
        cfg = workloads.cfg_example(N, R, wintype)
        assert [list(p[:3]) for p in cfg.params] == g["params"].tolist()
        outs, spec = make_gpu_chain(FDC, cfg).work_host(g["x"], want_spectrum=True)
        assert rel_l2(spec[:N], g["spectrum_first_block"]) < TOL and rel_l2(spec[-N:], g["spectrum_last_block"]) < TOL
        for i, o in enumerate(outs):
            assert rel_l2(o, g["out%d" % i]) < TOL, (name, i)


@pytest.mark.parametrize("N,ovl", [(1024, 768), (1024, 640), (32768, 24576)])
def test_sliding_window_overlap_above_half(FDC, N, ovl):
    """overlap > 50 % (SURVEY 8d 'true 75 % overlap'): the reference's overlap_save cannot express it
    (lib/overlap_save_impl.cc:74-78), the fp64 restatement with an explicit hop is the oracle"""
    from oracle import fdc_numpy as fnp
    cfg = workloads.ChanConfig("ovl_test", N, 4, workloads.example_channels(), workloads.HANN, ovl=ovl)
    nblocks = 11
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=5)
    want, wspec = fnp.channelize(x, cfg.N, cfg.R, cfg.params, cfg.windowtype, want_spectrum=True, ovl=ovl, shifts=cfg.shifts())
    g = make_gpu_chain(FDC, cfg)
    o1, s1 = g.work_host(x[:2 * cfg.hop], want_spectrum=True)        # fewer new samples than the overlap: history slides
    o2, s2 = g.work_host(x[2 * cfg.hop:], want_spectrum=True)
    assert rel_l2(np.concatenate([s1, s2]), wspec.reshape(-1)) < TOL
    for i in range(cfg.nchan):
        assert rel_l2(np.concatenate([o1[i], o2[i]]), want[i]) < TOL
    # device path, same stream in three calls
    import torch
    g2 = make_gpu_chain(FDC, cfg)
    d_in = torch.from_numpy(x.view(np.float32).copy()).cuda()
    got = [[] for _ in range(cfg.nchan)]
    pos = 0
    for nb in (1, 3, nblocks - 4):
        d_out = torch.empty(nb * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
        g2.work_device(d_in.data_ptr() + 8 * pos * cfg.hop, nb, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        a = d_out.cpu().numpy().view(np.complex64)
        for i, (off, ln) in enumerate(g2.out_slices(nb)):
            got[i].append(a[off:off + ln])
        pos += nb
    for i in range(cfg.nchan):
        assert rel_l2(np.concatenate(got[i]), want[i]) < TOL


@pytest.mark.parametrize("N,ovl,chunk", [(1024, 768, 1), (1024, 896, 4), (2048, 1920, 5)])
def test_overlap_above_half_with_chunks_shorter_than_the_history(FDC, N, ovl, chunk):
    """overlap > 50 % with chunks of fewer blocks than reach back into the previous call (nh = ceil(ovl / hop) up to 15):
    device path (history through the two-segment loads of the first chunks) and host path (chunk boundaries inside the head, pageable
    buffers through the staging slots) against the fp64 restatement"""
    import torch
    from oracle import fdc_numpy as fnp
    cfg = workloads.ChanConfig("ovl_test", N, 4, workloads.example_channels(), workloads.HANN, ovl=ovl)
    nblocks = 37
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=8)
    want, _ = fnp.channelize(x, cfg.N, cfg.R, cfg.params, cfg.windowtype, ovl=ovl, shifts=cfg.shifts())
    d_in = torch.from_numpy(x.view(np.float32).copy()).cuda()
    for path in ("device", "host"):
        g = make_gpu_chain(FDC, cfg)
        g.chunk_blocks = chunk
        got = [[] for _ in range(cfg.nchan)]
        pos = 0
        for nb in (2, 1, 19, nblocks - 22):
            if path == "device":
                d_out = torch.empty(nb * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
                g.work_device(d_in.data_ptr() + 8 * pos * cfg.hop, nb, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                a = d_out.cpu().numpy().view(np.complex64)
                outs = [a[off:off + ln] for off, ln in g.out_slices(nb)]
            else:
                outs, _ = g.work_host(x[pos * cfg.hop:(pos + nb) * cfg.hop])
            for i in range(cfg.nchan):
                got[i].append(np.array(outs[i]))
            pos += nb
        for i in range(cfg.nchan):
            assert rel_l2(np.concatenate(got[i]), want[i]) < TOL, (path, i)


def test_time_sharded_equals_single_stream(FDC):
    """SURVEY 8e on one GPU: the stream cut into per-rank runs (own halo, closed-form phase origin) gives bit-identical
    channel outputs to one context fed the whole stream"""
    from FDC import sharded
    cfg = workloads.cfg2()
    nblocks, world = 23, 4
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=9)
    whole, _ = make_gpu_chain(FDC, cfg).work_host(x)
    parts = [sharded.run_sharded(sharded.channelizer_worker(make_gpu_chain(FDC, cfg)), x, cfg.hop, cfg.ovl, nblocks, r, world)
             for r in range(world)]
    for i in range(cfg.nchan):
        got = np.concatenate([p[i] for p in parts if p is not None])
        assert np.array_equal(got.view(np.uint8), whole[i].view(np.uint8))


def test_slab_placement_builds_one_stream_ordered_buffer(FDC):
    """fdc_chan_work_device_slab: two "ranks" (contexts positioned with seek + halo) write their runs into ONE buffer laid
    out for the whole stream -- what the ranks do into the sink rank's peer memory (FDC.sharded.PeerSink); bit identical
    to a single context fed the whole stream"""
    import torch
    from FDC import sharded
    cfg = workloads.cfg2()
    per, world = 11, 3
    nblocks = per * world
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=19)
    whole, _ = make_gpu_chain(FDC, cfg).work_host(x)
    d_all = torch.zeros(nblocks * cfg.out_per_block * 2, dtype=torch.float32, device="cuda")
    for r in range(world):
        g = make_gpu_chain(FDC, cfg)
        halo, new = sharded.shard_input(x, cfg.hop, cfg.ovl, r * per, per)
        g.seek(r * per, halo)
        d_in = torch.from_numpy(np.ascontiguousarray(new).view(np.float32).copy()).cuda()
        g.work_device_slab(d_in.data_ptr(), per, d_all.data_ptr(), nblocks, r * per, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    got = d_all.cpu().numpy().view(np.complex64)
    off = 0
    for i in range(cfg.nchan):
        n = nblocks * cfg.params[i][2]
        assert np.array_equal(got[off:off + n].view(np.uint32), whole[i].view(np.uint32)), i
        off += n
    with pytest.raises(FDC.FDCError, match="outside the slab"):
        make_gpu_chain(FDC, cfg).work_device_slab(d_in.data_ptr(), per, d_all.data_ptr(), per, 1)


@pytest.mark.parametrize("case,world", [("cfg2", 3), ("example4096", 2), ("cfg1", 8)])
def test_channel_sharded_sinks_hold_their_channels_in_stream_order(FDC, case, world):
    """fdc_chan_set_sinks / fdc_chan_work_device_sinks (FDC.sharded.ChannelSinks on one GPU: `world` time-sharded contexts, `world`
    sink buffers, all local): after every rank has run its blocks, sink k holds exactly the channels it owns, complete and in
    stream order, bit-identical to one context fed the whole stream"""
    import torch
    from FDC import sharded
    cfg = {"cfg2": workloads.cfg2, "cfg1": workloads.cfg1, "example4096": lambda: workloads.cfg_example(4096, 4, workloads.HANN)}[case]()
    per = 7
    nblocks = per * world
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=29) + workloads.noise_input(nblocks * cfg.hop, 30) * np.float32(0.1)
    whole, _ = make_gpu_chain(FDC, cfg).work_host(x)
    louts = [p[2] for p in cfg.params]
    owner, per_sink = sharded.channel_owners(louts, world)
    assert owner == sorted(owner) and len(owner) == cfg.nchan                                            # contiguous runs
    sinks = [torch.full((max(1, nblocks * per_sink[k]) * 2,), float("nan"), dtype=torch.float32, device="cuda") for k in range(world)]
    for r in range(world):
        g = make_gpu_chain(FDC, cfg)
        halo, new = sharded.shard_input(x, cfg.hop, cfg.ovl, r * per, per)
        g.seek(r * per, halo)
        g.set_sinks([t.data_ptr() for t in sinks], owner, r)          # sink r is 'local': the others go through the staging slab + copies
        d_in = torch.from_numpy(np.ascontiguousarray(new).view(np.float32).copy()).cuda()
        g.work_device_sinks(d_in.data_ptr(), per, nblocks, r * per, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    for k in range(world):
        got = sinks[k].cpu().numpy().view(np.complex64)
        off = 0
        for i in [i for i, o in enumerate(owner) if o == k]:
            n = nblocks * louts[i]
            assert np.array_equal(got[off:off + n].view(np.uint32), whole[i].view(np.uint32)), (k, i)
            off += n
        assert off == nblocks * per_sink[k]
    with pytest.raises(FDC.FDCError, match="set_sinks first"):
        make_gpu_chain(FDC, cfg).work_device_sinks(d_in.data_ptr(), per, nblocks, 0)


def test_empty_and_ragged_calls(FDC, ref):
    """edge cases of the work() contract: zero items, a context without channels (spectrum only), calls of one block, and
    outputs independent of how the same stream is cut into calls (the reference keeps history and counters across calls)"""
    cfg = workloads.cfg_example(1024, 4, workloads.RAMP)
    x = workloads.tones_input(cfg, 7 * cfg.hop, seed=23)
    g = make_gpu_chain(FDC, cfg)
    o0, s0 = g.work_host(x[:0], want_spectrum=True)
    assert all(o.size == 0 for o in o0) and s0.size == 0 and g.blockcount == 0
    want, wspec = make_ref_chain(ref, cfg).run(x, nthreads=1, want_spectrum=True)
    parts = [g.work_host(x[a * cfg.hop:b * cfg.hop], want_spectrum=True) for a, b in ((0, 1), (1, 2), (2, 2), (2, 7))]
    for i in range(cfg.nchan):
        assert rel_l2(np.concatenate([p[0][i] for p in parts]), want[i]) < TOL
    assert rel_l2(np.concatenate([p[1] for p in parts]), wspec) < TOL
    # no channels: the front end alone (debug spectrum port)
    g0 = FDC.Channelizer(cfg.N, cfg.ovl, cfg.R, [])
    outs, spec = g0.work_host(x, want_spectrum=True)
    assert outs == [] and rel_l2(spec, wspec) < TOL
    # constructor errors of the C ABI surface as FDCError with the reason
    with pytest.raises(FDC.FDCError, match="power of two"):
        FDC.Channelizer(1000, 250, 4, [])
    with pytest.raises(FDC.FDCError, match="outside the spectrum"):
        FDC.Channelizer(1024, 256, 4, [(1000, 64, 48, 0, 64.0, FDC.psw_tables(64, 4, 0.5, 0.75, 1))])


def test_context_follows_its_device_across_threads(FDC):
    """CUDA's current device is per thread and GNU Radio calls work() from a scheduler thread: a context made on device 1 must
    work when called from a thread whose current device is 0 (needs two GPUs; the single-GPU box still runs the thread part)"""
    import threading
    from oracle import fdc_numpy as fnp
    L = FDC._cabi.lib()
    dev = 1 if L.fdc_device_count() > 1 else 0
    FDC._cabi.check(L.fdc_set_device(dev))
    try:
        cfg = workloads.cfg_example(1024, 4, workloads.HANN)
        chan = make_gpu_chain(FDC, cfg)
        sd = FDC.SegmentDetection(1, 1024, 4, 0.1, 0.9, 10.0, 0.0312, 0.2, 4, 1, True, False, "", False, 0)
    finally:
        FDC._cabi.check(L.fdc_set_device(0))
    x = workloads.tones_input(cfg, 40 * cfg.hop, seed=77)
    spec, _ = sc.bursty_spectra(1024, 40, 6, seed=3, widths=(16, 32), raster=64)
    box = {}

    def worker():                                   # a fresh thread: current device 0
        try:
            box["outs"], _ = chan.work_host(x)
            sd.work(40, [spec.reshape(-1)])
            box["msgs"] = sd.messages()
        except Exception as e:                      # noqa: BLE001
            box["err"] = e
    t = threading.Thread(target=worker); t.start(); t.join()
    assert "err" not in box, box.get("err")
    want, _ = fnp.channelize(x, cfg.N, cfg.R, cfg.params, cfg.windowtype)
    for o, w in zip(box["outs"], want):
        assert rel_l2(o, w) < TOL
    assert len(box["msgs"]) >= 1
