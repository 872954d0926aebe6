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
