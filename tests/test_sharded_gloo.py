"""CPU tests of the multi-GPU plumbing: time sharding with per-rank halo, closed-form phase origin and the
gather of the per-channel runs to the sink rank, on torch.distributed/gloo with world_size 2 and 3.  The per-rank worker
here is the fp64 oracle (allowed in tests/); on the GPU box the same functions drive FDC.Channelizer
(test_gpu_chan.py::test_time_sharded_equals_single_stream)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _oracle_worker(cfg):
    from oracle import fdc_numpy as fnp

    def work(halo, new, first_block):
        # counter of phase_shifting_windowing_vcc after first_block blocks: (first_block * shift) mod R
        c0 = [((first_block % cfg.R) * (((p[0] % cfg.R) + cfg.R) % cfg.R)) % cfg.R for p in cfg.params]
        outs, _ = fnp.channelize(new, cfg.N, cfg.R, cfg.params, cfg.windowtype, hist=np.asarray(halo, dtype=np.complex128), counter0=c0)
        return [o.astype(np.complex64) for o in outs]
    return work


def _rank_main(rank, world, port, nblocks, q):
    for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import workloads
    from FDC import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = workloads.cfg_example(1024, 4, workloads.HANN)
        x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=31)
        louts = [p[2] for p in cfg.params]
        local = sharded.run_sharded(_oracle_worker(cfg), x, cfg.hop, cfg.ovl, nblocks, rank, world)
        outs = sharded.gather_outputs(local, louts, nblocks, rank, world, dst=0)
        meta = sharded.gather_objects({"rank": rank, "blocks": sharded.partition(nblocks, world)[1][rank]}, rank, world, dst=0)
        if rank == 0:
            q.put((outs, meta))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nblocks", [(2, 9), (3, 7), (2, 1)])
def test_time_sharding_and_gather(world, nblocks):
    import torch.multiprocessing as mp
    import workloads
    from oracle import fdc_numpy as fnp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, nblocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs, meta = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    cfg = workloads.cfg_example(1024, 4, workloads.HANN)
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=31)
    want, _ = fnp.channelize(x, cfg.N, cfg.R, cfg.params, cfg.windowtype)
    assert [m["rank"] for m in meta] == list(range(world))
    assert sum(m["blocks"] for m in meta) == nblocks
    for i, w in enumerate(want):
        assert outs[i].size == w.size
        # identical arithmetic on identical samples: only the fp32 rounding of the gathered slabs differs
        assert np.max(np.abs(outs[i] - w)) <= 1e-6 * max(1.0, float(np.max(np.abs(w))))


def test_partition_and_halo():
    for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from FDC import sharded
    assert sharded.partition(10, 4) == ([0, 3, 6, 8], [3, 3, 2, 2])
    assert sharded.partition(2, 4) == ([0, 1, 2, 2], [1, 1, 0, 0])
    x = np.arange(100, dtype=np.complex64)
    halo, new = sharded.shard_input(x, 10, 4, 0, 3)
    assert np.all(halo == 0) and halo.size == 4 and np.array_equal(new, x[:30])
    halo, new = sharded.shard_input(x, 10, 4, 3, 2)
    assert np.array_equal(halo, x[26:30]) and np.array_equal(new, x[30:50])
    # overlap larger than the hop: the halo of an early block straddles the stream history
    hist = -np.arange(1, 26, dtype=np.complex64)
    halo, new = sharded.shard_input(x, 10, 25, 1, 2, stream_history=hist)
    assert np.array_equal(halo, np.concatenate([hist[10:], x[:10]])) and np.array_equal(new, x[10:30])


# ------------------------------------------------------------------------------------------------ activity-gated blocks, time sharded
def _activity_blocks(FDC, N):
    """(name, block, power rows) for the three activity-gated blocks on one bursty scenario (host-logic contexts: no GPU)"""
    import scenarios as sc
    x, _ = sc.bursty_spectra(N, 96, 8, seed=52, widths=(16, 32, 64), mean_on=9, mean_off=12)
    res = []
    b = FDC.SegmentDetection(5, N, 4, 0.1, 0.9, 10.0, 0.0312, 0.2, 4, 1, True, False, "", False, 0, _logic=True)
    st = b.state()
    res.append(("segdet", b, sc.group_power(x, st["d_start"], st["D"], st["M"])))
    b = FDC.activity_detection_channelizer_vcm(N, [[0.1, 0.45], [0.55, 0.9]], 10.0, 4, 4, True, False, "", False, 0.0312, 1, 0.2, 0, _logic=True)
    res.append(("actdet", b, np.concatenate([sc.group_power(x, s["start"], s["D"], s["M"], mean=True) for s in b.segments()], axis=1)))
    b = FDC.PowerActivationChannel(N, 0.45, 0.1, 4, 6.0, 3, 1, True, False, "", 0, 7, _logic=True)
    st = b.state()
    # a carrier of its own inside the measured band so that the channel really toggles
    y = x.copy(); y[:, st["measure_start"]:st["measure_stop"]] *= np.where((np.arange(96) // 7) % 2, 8.0, 1.0).astype(np.float32)[:, None]
    res.append(("pac", b, sc.band_power(y, st["measure_start"], st["measure_stop"])))
    return res


ACT_CALLS = (40, 1, 55)       # global calls (blocks each): 96 blocks in all


def _rank_activity(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import FDC
    import scenarios as sc
    from FDC import sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got = {}
        for name, blk, P in _activity_blocks(FDC, 1024):
            sa = sharded.ShardedActivity(blk, rank, world, dst=0)
            msgs, pos = [], 0
            for n in ACT_CALLS:
                first, count = sharded.partition(n, world)
                m = sa.work(n, power=P[pos + first[rank]:pos + first[rank] + count[rank]])
                pos += n
                if rank == 0:
                    msgs += m
                else:
                    assert m is None
            got[name] = [sc.meta_tuple(m) for m in msgs]
        # the three blocks as one group (one all-gather + one gather per call), local phases on a thread pool
        from concurrent.futures import ThreadPoolExecutor
        trio = _activity_blocks(FDC, 1024)
        grp = sharded.ShardedActivityGroup([t[1] for t in trio], rank, world, dst=0, pool=ThreadPoolExecutor(3))
        msgs, pos = [[] for _ in trio], 0
        for n in ACT_CALLS:
            first, count = sharded.partition(n, world)
            m = grp.work(n, powers=[t[2][pos + first[rank]:pos + first[rank] + count[rank]] for t in trio])
            pos += n
            if rank == 0:
                for i, mm in enumerate(m):
                    msgs[i] += mm
        for t, mm in zip(trio, msgs):
            got["group_" + t[0]] = [sc.meta_tuple(m) for m in mm]
        if rank == 0:
            q.put(got)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_activity_bookkeeping(world):
    """measure (own rows) -> all-gather of the records -> replicated bookkeeping -> extract (own jobs) -> assemble on the sink:
    the PDU sequence equals the single-stream block's (which test_host_logic.py checks against the reference blocks)"""
    import torch.multiprocessing as mp
    for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import FDC
    import scenarios as sc
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_activity, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for name, blk, P in _activity_blocks(FDC, 1024):
        blk.logic_work(P.shape[0], P)
        want = [sc.meta_tuple(m) for m in blk.messages()]
        assert len(want) >= 3, name
        assert got[name] == want, name
        assert got["group_" + name] == want, name


def test_shard_layout_is_the_same_on_every_rank_and_covers_the_call():
    """fdc_*_shard_layout (device form of the time-sharded call): every rank derives the offsets of ALL jobs of the call from
    the replicated job list; the run holds exactly the samples of all ranks, by channel or in job order"""
    for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import FDC
    from FDC import sharded
    for name, mk in (("segdet", 0), ("actdet", 1), ("pac", 2)):
        ranks = [_activity_blocks(FDC, 1024)[mk] for _ in range(3)]                # three "ranks": the same block built three times
        P = ranks[0][2]
        n = P.shape[0]
        first, count = sharded.partition(n, 3)
        recs = b"".join(ranks[r][1].shard_measure(count[r], power=P[first[r]:first[r] + count[r]]) for r in range(3))
        with pytest.raises(RuntimeError):
            ranks[0][1].shard_layout(True)                                          # nothing decided yet
        njobs = [ranks[r][1].shard_decide(n, recs) for r in range(3)]
        assert len(set(njobs)) == 1 and njobs[0] > 0, name
        sizes = [ranks[0][1].shard_samples(first[r], count[r]) for r in range(3)]
        for by_channel in (True, False):
            totals = [ranks[r][1].shard_layout(by_channel) for r in range(3)]
            assert totals == [sum(sizes)] * 3, (name, by_channel)
        for r in range(3):
            ranks[r][1].shard_assemble(None)


def test_channel_owners_partition():
    """FDC.sharded.channel_owners: contiguous runs, every channel owned, volumes balanced to within one channel"""
    from FDC import sharded
    rng = np.random.default_rng(3)
    for trial in range(200):
        n = int(rng.integers(1, 300)); world = int(rng.integers(1, 9))
        louts = [int(v) for v in rng.choice([24, 48, 96, 192, 384, 768], size=n)]
        owner, per_sink = sharded.channel_owners(louts, world)
        assert len(owner) == n and owner == sorted(owner) and owner[0] >= 0 and owner[-1] < world
        assert per_sink == [sum(lo for lo, o in zip(louts, owner) if o == k) for k in range(world)]
        assert sum(per_sink) == sum(louts)
        if n >= 4 * world:
            assert max(per_sink) - min(per_sink) <= 2 * max(louts), (louts, world, per_sink)
    assert sharded.channel_owners([384] * 256, 8)[1] == [12288] * 8
