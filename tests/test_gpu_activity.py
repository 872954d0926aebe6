"""GPU parity of the activity-gated blocks (PowerActivationChannel, SegmentDetection, activity_detection_channelizer_vcm)
and of the hier block mirror against the oracle.  Integers (geometry, block indices, vector ranges, part numbers, sample
counts, PDU order) bit exact; payload samples rel-L2 <= 1e-5; decimated powers bit exact (same summation order)."""
import numpy as np
import pytest

import geometry
import scenarios as sc
import workloads
from helpers import rel_l2, make_ref_chain

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def FDC():
    import FDC as m
    return m


def feed(block, x, N, chunks):
    pos = 0
    for n in chunks:
        block.work(n, [x[pos * N:(pos + n) * N]]) if hasattr(block, "_h") else block.work(x[pos * N:(pos + n) * N])
        pos += n
    assert pos * N == x.size


def compare_messages(ma, mb, ordered=True):
    ka = [sc.meta_tuple(m) for m in ma]; kb = [sc.meta_tuple(m) for m in mb]
    if ordered:
        assert ka == kb
        pairs = zip(ma, mb)
    else:
        assert sorted(ka) == sorted(kb)
        ib = {k: m for k, m in zip(kb, mb)}
        pairs = [(m, ib[k]) for k, m in zip(ka, ma)]
    worst = 0.0
    for a, b in pairs:
        assert a["data"].size == b["data"].size
        if a["data"].size:
            worst = max(worst, rel_l2(b["data"], a["data"]))
    assert worst < TOL, worst


def test_pac_b8(FDC, ref):
    x = sc.b8_input().reshape(-1)
    args = (256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7)
    a = ref.PowerActivationChannel(*args); b = FDC.PowerActivationChannel(*args)
    feed(a, x, 256, (5, 3, 8)); feed(b, x, 256, (1, 9, 6))
    assert a.state() == b.state()
    assert np.array_equal(a.tables().view(np.uint32), b.tables().view(np.uint32))
    ma, mb = a.messages(), b.messages()
    assert [(m["part"], m["blockstart"], m["blockend"], m["data"].size) for m in mb] == [(0, 3, 6, 72), (1, 3, 9, 72), (2, 3, 11, 48)]
    compare_messages(ma, mb)


@pytest.mark.parametrize("maxblocks", [-1, 0, 5])
def test_pac_bursty(FDC, ref, maxblocks):
    N = 4096
    x, truth = sc.bursty_spectra(N, 150, 6, seed=4, widths=(64, 128), raster=512, mean_on=12, mean_off=15)
    start, stop = truth[0][0], truth[0][1]
    cf = (start + stop) / 2.0 / N; bw = (stop - start) / float(N)
    args = (N, cf, bw, 4, 6.0, maxblocks, 0, True, False, "", 0, 2)
    a = ref.PowerActivationChannel(*args); b = FDC.PowerActivationChannel(*args)
    feed(a, x.reshape(-1), N, (150,)); feed(b, x.reshape(-1), N, (40, 1, 64, 45))
    ma, mb = a.messages(), b.messages()
    assert len(ma) >= 2
    compare_messages(ma, mb)
    sa, sb = a.state(), b.state()
    assert sa == sb


def test_segdet_b9(FDC, ref):
    x = sc.b9_input().reshape(-1)
    args = (3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, True, False, "", False, 0)
    a = ref.SegmentDetection(*args); b = FDC.SegmentDetection(*args)
    feed(a, x, 256, (3, 4, 2, 11)); feed(b, x, 256, (8, 1, 11))
    mb = b.messages()
    got = [(sc.strip_time(m["ID"]), m["finalized"], m["part"], m["blockstart"], m["blockend"], m["vectorstart"], m["vectorend"], m["data"].size) for m in mb]
    assert got == [("DETECTED.3.0", False, 0, 2, 6, 80, 144, 192), ("DETECTED.3.1", False, 0, 4, 8, 132, 196, 192),
                   ("DETECTED.3.0", False, 1, 2, 10, 80, 144, 192), ("DETECTED.3.0", True, 2, 3, 11, 80, 144, 0),
                   ("DETECTED.3.1", False, 1, 4, 12, 132, 196, 192), ("DETECTED.3.1", True, 2, 5, 16, 132, 196, 144)]
    compare_messages(a.messages(), mb)
    assert np.array_equal(a.power().view(np.uint32), b.power().view(np.uint32))
    assert a.active_channels() == b.active_channels()


@pytest.mark.parametrize("N,maxblocks,delay,seed", [(4096, 8, 1, 1), (4096, -1, 0, 2), (16384, 0, 2, 3), (1024, 3, 1, 4)])
def test_segdet_bursty(FDC, ref, N, maxblocks, delay, seed):
    nblocks = 100
    x, truth = sc.bursty_spectra(N, nblocks, 12, seed=seed, widths=(16, 32, 64, 128) if N > 1024 else (16, 32), raster=256 if N > 1024 else 64)
    args = (1, N, 4, 0.1, 0.9, 10.0, 0.002 * (4096.0 / N) * 2, 0.2, maxblocks, delay, True, False, "", False, 0)
    a = ref.SegmentDetection(*args); b = FDC.SegmentDetection(*args)
    feed(a, x.reshape(-1), N, (nblocks,)); feed(b, x.reshape(-1), N, (33, 1, 2, 64))
    ma, mb = a.messages(), b.messages()
    assert len(ma) >= 5
    compare_messages(ma, mb)
    assert a.active_channels() == b.active_channels()
    assert np.array_equal(a.power().view(np.uint32), b.power().view(np.uint32))
    sa, sb = a.state(), b.state()
    assert sa == sb


def test_segdet_noise_only_many_edges(FDC, ref):
    """D = 1 on pure noise: hundreds of edges per block, more than the compact device lists hold -> host fallback path"""
    N = 4096
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((12, N)) + 1j * rng.standard_normal((12, N))).astype(np.complex64)
    args = (0, N, 4, 0.02, 0.999, 3.0, 0.0001, 0.1, 2, 0, True, False, "", False, 0)
    a = ref.SegmentDetection(*args); b = FDC.SegmentDetection(*args)
    assert a.state()["D"] == 1
    feed(a, x.reshape(-1), N, (12,)); feed(b, x.reshape(-1), N, (5, 7))
    ma, mb = a.messages(), b.messages()
    assert len(ma) > 50
    compare_messages(ma, mb)


@pytest.mark.parametrize("threads", [False, True])
def test_actdet_b11_and_bursty(FDC, ref, threads):
    x = sc.b9_input().reshape(-1)
    args = (256, [[0.1, 0.9]], 10.0, 4, 4, True, False, "", threads, 0.0625, 1, 0.2, 0)
    a = ref.activity_detection_channelizer_vcm(*args); b = FDC.activity_detection_channelizer_vcm(*args)
    feed(a, x, 256, (3, 4, 2, 11)); feed(b, x, 256, (10, 10))
    mb = b.messages()
    got = sorted((sc.strip_time(m["ID"]), m["finalized"], m["part"], m["blockstart"], m["blockend"], m["data"].size) for m in mb)
    assert got == sorted([("DETECTED.0.0", False, 0, 3, 7, 192), ("DETECTED.0.1", False, 0, 5, 9, 192), ("DETECTED.0.0", False, 1, 3, 11, 192),
                          ("DETECTED.0.0", True, 2, 4, 12, 0), ("DETECTED.0.1", False, 1, 5, 13, 192), ("DETECTED.0.1", True, 2, 6, 17, 144)])
    compare_messages(a.messages(), mb, ordered=not threads)
    N = 4096
    x, truth = sc.bursty_spectra(N, 80, 14, seed=8, widths=(32, 64), raster=256)
    args = (N, [[0.1, 0.45], [0.55, 0.9]], 10.0, 4, 6, True, False, "", threads, 0.004, 1, 0.2, 0)
    a = ref.activity_detection_channelizer_vcm(*args); b = FDC.activity_detection_channelizer_vcm(*args)
    feed(a, x.reshape(-1), N, (80,)); feed(b, x.reshape(-1), N, (17, 63))
    ma, mb = a.messages(), b.messages()
    assert len(ma) >= 5
    compare_messages(ma, mb, ordered=not threads)
    assert a.segments() == b.segments()
    for i in range(2):
        assert np.array_equal(a.power(i).view(np.uint32), b.power(i).view(np.uint32))


def test_file_output(FDC, ref, tmp_path):
    x = sc.b9_input().reshape(-1)
    pa = tmp_path / "a"; pb = tmp_path / "b"; pa.mkdir(); pb.mkdir()
    a = ref.SegmentDetection(3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, False, True, str(pa), False, 0)
    b = FDC.SegmentDetection(3, 256, 4, 0.1, 0.9, 10.0, 0.0625, 0.2, 4, 1, False, True, str(pb), False, 0)
    feed(a, x, 256, (20,)); feed(b, x, 256, (20,))
    assert a.messages() == [] and b.messages() == []
    fa = sorted(p.name.split(".", 1)[1] for p in pa.iterdir()); fb = sorted(p.name.split(".", 1)[1] for p in pb.iterdir())
    assert fa == fb and len(fa) == 6
    for na in pa.iterdir():
        nb = [p for p in pb.iterdir() if p.name.split(".", 1)[1] == na.name.split(".", 1)[1]][0]
        da = np.fromfile(str(na), dtype=np.complex64); db = np.fromfile(str(nb), dtype=np.complex64)
        assert da.size == db.size
        if da.size:
            assert rel_l2(db, da) < TOL


def test_hier_block_against_reference_flowgraph(FDC, ref):
    """FrequencyDomainChannelizer mirror (25 GRC arguments) vs the same flowgraph assembled from the reference blocks:
    throughput channels + PowerActivationChannel + SegmentDetection + debug spectrum, FDC_example.grc parameters."""
    N, R = 4096, 4
    chans = workloads.example_channels()
    blk = FDC.FrequencyDomainChannelizer(8, 1, N, R, chans, chans, 4.0, 32000.0, 0.0, 'normalized', 0, True, False, "", True,
                                         [[-0.4, 0.4]], 10.0, 0.01, 1, 0.2, 0, 1, 8, 8, True)
    cfg = workloads.cfg_example(N, R, workloads.RECTANGULAR)
    assert [p[:3] for p in blk.channel_params] == [p[:3] for p in cfg.params]
    hop = cfg.hop; nblocks = 60
    rng = np.random.default_rng(2)
    n = np.arange(nblocks * hop)
    x = 0.05 * (rng.standard_normal(n.size) + 1j * rng.standard_normal(n.size))
    for i, (fq, bw) in enumerate(chans):                       # gated tones: on during different block ranges
        gate = ((n // hop) % 20 >= 3 * i + 2) & ((n // hop) % 20 < 3 * i + 9)
        x = x + gate * np.exp(2j * np.pi * (fq + 0.1 * bw) * n)
    x = x.astype(np.complex64)
    outs1 = blk.work(x[:25 * hop]); outs2 = blk.work(x[25 * hop:])
    want, wspec = make_ref_chain(ref, cfg).run(x, nthreads=4, want_spectrum=True)
    spec = np.concatenate([outs1[0].reshape(-1), outs2[0].reshape(-1)])
    assert rel_l2(spec, wspec) < TOL
    for i in range(len(chans)):
        assert rel_l2(np.concatenate([outs1[1 + i], outs2[1 + i]]), want[i]) < TOL
    ref_msgs = []
    for i, (fq, bw) in enumerate(chans):
        p = ref.PowerActivationChannel(N, geometry.get_freq(fq), geometry.get_bw(bw), R, 4.0, 8, 1, True, False, "", 0, i)
        p.work(wspec); ref_msgs += p.messages()
    s = ref.SegmentDetection(0, N, R, geometry.get_freq(-0.4), geometry.get_freq(0.4), 10.0, geometry.get_bw(0.01), 0.2, 8, 1, True, False, "", True, 0)
    s.work(wspec); ref_msgs += s.messages()
    got = blk.messages()
    assert len(got) == len(ref_msgs) and len(got) > 8
    compare_messages(ref_msgs, got, ordered=False)


@pytest.mark.parametrize("freqmode", ["basebandfs", "centerfreqfs"])
def test_hier_block_frequency_modes(FDC, ref, freqmode):
    """the hier block in basebandfs / centerfreqfs mode (python/FrequencyDomainChannelizer.py:70-91,
    322-345), channel and segment frequencies given in Hz.  Geometry against tests/golden/freqmodes.json (the reference's own
    lines, executed); samples and PDUs against the reference blocks built from the golden normalised values."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "freqmodes.json")) as fh:
        case = [c for c in json.load(fh)["cases"] if c["freqmode"] == freqmode and c["blocksize"] == 4096][0]
    N, R = case["blocksize"], case["relinvovl"]
    fs, cf = case["fs"], case["centerfrequency"]
    chans = [tuple(c) for c in case["channels"]]
    segs = [list(sg) for sg in case["segments"]][:1]
    blk = FDC.FrequencyDomainChannelizer(8, 1, N, R, chans, chans[:2], 4.0, fs, cf, freqmode, workloads.HANN, True, False, "", True,
                                         segs, 10.0, 0.01 * fs, 1, 0.2, 0, 1, 8, 8, True)
    assert [list(p) for p in blk.channel_params] == case["channel_params"]
    assert blk.throughput_channels == case["normalized_channels"]
    assert blk.activity_detection_segments == case["normalized_segments"][:1]
    params = [tuple(p) for p in case["channel_params"]]
    hop = N - N // R; nblocks = 48
    rng = np.random.default_rng(5)
    n = np.arange(nblocks * hop)
    x = 0.05 * (rng.standard_normal(n.size) + 1j * rng.standard_normal(n.size))
    for i, (fq, bw) in enumerate(case["normalized_channels"]):   # gated tones at the channel centres (normalised, 0 .. 1 over the shifted band)
        gate = ((n // hop) % 16 >= 2 * i + 1) & ((n // hop) % 16 < 2 * i + 8)
        x = x + gate * np.exp(2j * np.pi * (fq - 0.5 + 0.1 * bw) * n)
    x = x.astype(np.complex64)
    outs1 = blk.work(x[:20 * hop]); outs2 = blk.work(x[20 * hop:])
    want, wspec = ref.Chain(N, R, params, workloads.HANN).run(x, nthreads=4, want_spectrum=True)
    assert rel_l2(np.concatenate([outs1[0].reshape(-1), outs2[0].reshape(-1)]), wspec) < TOL
    for i in range(len(chans)):
        assert rel_l2(np.concatenate([outs1[1 + i], outs2[1 + i]]), want[i]) < TOL
    ref_msgs = []
    for i, (fq, bw) in enumerate(case["normalized_channels"][:2]):
        p = ref.PowerActivationChannel(N, fq, bw, R, 4.0, 8, 1, True, False, "", 0, i)
        p.work(wspec); ref_msgs += p.messages()
    a, b = case["normalized_segments"][0]
    s = ref.SegmentDetection(0, N, R, a, b, 10.0, (0.01 * fs / fs) % 1.0, 0.2, 8, 1, True, False, "", True, 0)
    s.work(wspec); ref_msgs += s.messages()
    got = blk.messages()
    assert len(got) == len(ref_msgs) and len(got) > 4
    compare_messages(ref_msgs, got, ordered=False)


def test_hier_block_transformed_input_mode(FDC, ref):
    """inpveclen = blocksize (python/FrequencyDomainChannelizer.py:284-290): the input items are fft-shifted, unnormalised
    spectra; oracle = the reference blocks chained by hand behind the restated multiply_const / inverse fft_vcc stages"""
    N, R = 2048, 4
    chans = [(0.12, 0.05), (-0.2, 0.02), (0.31, 0.1)]
    rng = np.random.default_rng(17)
    nblocks = 9
    X = (rng.standard_normal((nblocks, N)) + 1j * rng.standard_normal((nblocks, N))).astype(np.complex64) * np.float32(30.0)
    blk = FDC.FrequencyDomainChannelizer(8, N, N, R, chans, [], 6.0, 1.0, 0.0, "normalized", 1, False, False, "", False, [], 10.0, 0.01,
                                         1, 0.2, 0, 1, 4, 4, True)
    o1 = blk.work(X[:4].reshape(-1)); o2 = blk.work(X[4:].reshape(-1))
    spec = np.concatenate([o1[0], o2[0]])
    want_spec = (X * np.float32(1.0 / N)).astype(np.complex64)                    # multiply_const_cc(1/N): exact (power of two)
    assert np.array_equal(spec.view(np.uint32), want_spec.view(np.uint32))
    for i, (fq, bw) in enumerate(chans):
        f, l, lout, pb, sb = geometry.get_opt_channelparams(N, R, geometry.get_freq(fq), geometry.get_bw(bw))
        cut0 = ref.vector_cut_vxx(8, N, f, l); psw = ref.phase_shifting_windowing_vcc(l, R, f, pb, sb, 1); cut3 = ref.vector_cut_vxx(8, l, l - lout, lout)
        y = cut0.work(want_spec.reshape(-1)).view(np.complex64)
        y = psw.work(y).view(np.complex64)
        y = ref.fft_vcc(y, l, False, True)
        y = cut3.work(y).view(np.complex64) * np.float32(l)
        got = np.concatenate([o1[1 + i], o2[1 + i]])
        assert got.size == y.size == nblocks * lout
        assert rel_l2(got, y) < 1e-5


# ------------------------------------------------------------------------------------------------ time-sharded calls
def _virtual_ranks_run(mk, x, N, calls, world, device_form=False):
    """The fdc_*_shard_* sequence of FDC/sharded.py: ShardedActivity with `world` contexts in ONE process (no process group
    needed to check the arithmetic): every virtual rank holds only its own rows plus the one before them on the device."""
    import torch
    from FDC import sharded
    ranks = [mk() for _ in range(world)]
    rows = x.reshape(-1, N)
    msgs, pos = [], 0
    for n in calls:
        first, count = sharded.partition(n, world)
        dev, prev = [], []
        for r in range(world):
            lo = pos + first[r]
            own = torch.from_numpy(np.ascontiguousarray(rows[lo:lo + count[r]]).view(np.float32)).cuda()
            before = torch.from_numpy(np.ascontiguousarray(rows[lo - 1:lo]).view(np.float32)).cuda() if lo > 0 else None
            dev.append(own); prev.append(before)
        recs = [ranks[r].shard_measure(count[r], dev[r].data_ptr()) for r in range(world)]
        blob = b"".join(recs)
        njobs = [ranks[r].shard_decide(n, blob) for r in range(world)]
        assert len(set(njobs)) == 1
        if device_form:          # the gather fused into the extract kernel: every rank stores into the sink's device buffer
            sizes = [ranks[0].shard_samples(first[r], count[r]) for r in range(world)]
            sink = torch.zeros(2 * max(sum(sizes), 1), dtype=torch.float32, device="cuda")
            for r in range(world):
                if device_form == "by_channel":     # one run for all ranks, channel by channel: every rank stores at the run's start + its offsets
                    assert ranks[r].shard_layout(True) == sum(sizes)
                got = ranks[r].shard_extract_device(first[r], count[r], dev[r].data_ptr(), prev[r].data_ptr() if prev[r] is not None else 0,
                                                    sink.data_ptr() + (0 if device_form == "by_channel" else 8 * sum(sizes[:r])))
                assert got == sizes[r]
            for r in range(1, world):
                ranks[r].shard_assemble(None)
            ranks[0].shard_assemble_device(sink.data_ptr(), sum(sizes))
        else:
            parts = [ranks[r].shard_extract(first[r], count[r], dev[r].data_ptr(), prev[r].data_ptr() if prev[r] is not None else 0)
                     for r in range(world)]
            for r in range(1, world):
                ranks[r].shard_assemble(None)
            ranks[0].shard_assemble(np.concatenate(parts))
        msgs += ranks[0].messages()
        pos += n
    return msgs, ranks


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_segdet_equals_single_stream(FDC, ref, world):
    N, nblocks = 4096, 100
    x, _ = sc.bursty_spectra(N, nblocks, 12, seed=5, widths=(16, 32, 64, 128), raster=256)
    args = (1, N, 4, 0.1, 0.9, 10.0, 0.004, 0.2, 8, 1, True, False, "", False, 0)
    a = ref.SegmentDetection(*args); feed(a, x.reshape(-1), N, (nblocks,))
    b = FDC.SegmentDetection(*args); feed(b, x.reshape(-1), N, (nblocks,))
    mb = b.messages()
    ms, ranks = _virtual_ranks_run(lambda: FDC.SegmentDetection(*args), x, N, (37, 1, 62), world)
    assert len(mb) >= 5
    compare_messages(a.messages(), ms)
    # against the single-stream GPU block: the same kernels on the same rows -> identical bits
    assert [sc.meta_tuple(m) for m in mb] == [sc.meta_tuple(m) for m in ms]
    for u, v in zip(mb, ms):
        assert np.array_equal(u["data"].view(np.uint32), v["data"].view(np.uint32))
    assert all(r.active_channels() == b.active_channels() for r in ranks)
    for form in (True, "by_channel"):
        md, _ = _virtual_ranks_run(lambda: FDC.SegmentDetection(*args), x, N, (37, 1, 62), world, device_form=form)
        assert [sc.meta_tuple(m) for m in mb] == [sc.meta_tuple(m) for m in md]
        for u, v in zip(mb, md):
            assert np.array_equal(u["data"].view(np.uint32), v["data"].view(np.uint32))


def test_sharded_pac_and_actdet_equal_single_stream(FDC, ref):
    N, nblocks = 4096, 80
    x, _ = sc.bursty_spectra(N, nblocks, 14, seed=8, widths=(32, 64), raster=256)
    args = (N, [[0.1, 0.45], [0.55, 0.9]], 10.0, 4, 6, True, False, "", False, 0.004, 1, 0.2, 0)
    a = ref.activity_detection_channelizer_vcm(*args); feed(a, x.reshape(-1), N, (nblocks,))
    ms, _ = _virtual_ranks_run(lambda: FDC.activity_detection_channelizer_vcm(*args), x, N, (50, 30), 3)
    assert len(ms) >= 5
    compare_messages(a.messages(), ms)
    x = sc.b8_input().reshape(-1)
    args = (256, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7)
    a = ref.PowerActivationChannel(*args); feed(a, x, 256, (16,))
    ms, ranks = _virtual_ranks_run(lambda: FDC.PowerActivationChannel(*args), x, 256, (7, 9), 4)
    assert [(m["part"], m["blockstart"], m["blockend"], m["data"].size) for m in ms] == [(0, 3, 6, 72), (1, 3, 9, 72), (2, 3, 11, 48)]
    compare_messages(a.messages(), ms)
    assert all(r.state() == a.state() for r in ranks)


def test_segdet_cfg5_geometry(FDC, ref):
    """configs[4] class: FFT 262144, narrow (40-bin) DAMA carriers on a 64-bin raster over [0.02, 0.98], hundreds of them active per
    block; single stream and time sharded over 4 virtual ranks against the reference block"""
    N, nblocks = 262144, 14
    x, truth = sc.bursty_spectra(N, nblocks, 1500, seed=55, raster=64, widths=(40,), mean_on=4, mean_off=30, lo=0.02, hi=0.98)
    args = (2, N, 4, 0.02, 0.98, 10.0, 16.0 / N, 0.2, 3, 1, True, False, "", False, 0)
    a = ref.SegmentDetection(*args); b = FDC.SegmentDetection(*args)
    assert a.state() == b.state() and a.state()["D"] <= 16
    feed(a, x.reshape(-1), N, (nblocks,)); feed(b, x.reshape(-1), N, (5, 9))
    ma, mb = a.messages(), b.messages()
    assert len(ma) >= 300
    compare_messages(ma, mb)
    assert a.active_channels() == b.active_channels()
    assert np.array_equal(a.power().view(np.uint32), b.power().view(np.uint32))
    ms, _ = _virtual_ranks_run(lambda: FDC.SegmentDetection(*args), x, N, (nblocks,), 4, device_form="by_channel")
    assert [sc.meta_tuple(m) for m in mb] == [sc.meta_tuple(m) for m in ms]
    for u, v in zip(mb, ms):
        assert np.array_equal(u["data"].view(np.uint32), v["data"].view(np.uint32))
