"""Multi-GPU correctness check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multirank_check.py

1. channel-sharded sinks (FDC.sharded.ChannelSinks): the time-sharded ranks' extract kernels store every channel's rows into its
   owner's buffer over NVLink peer memory; afterwards rank k holds its channels complete and in stream order, bit-identical to one
   context fed the whole stream.
2. instance-sharded activity sinks (ShardedActivityGroup with PeerBuffers): SegmentDetection and PowerActivationChannel instances
   on a time-sharded spectrum; every PDU (metadata and samples) equals the single-stream block's, published on the owning rank.
tests/test_gpu_multirank.py launches this when the box has at least two GPUs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-fdc_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    import FDC
    import scenarios as sc
    import workloads
    from FDC import sharded
    from helpers import make_gpu_chain
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    FDC._cabi.check(FDC._cabi.lib().fdc_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- 1. channel-sharded sinks ----
    cfg = workloads.cfg2()
    per = 9; nblocks = per * world
    x = workloads.tones_input(cfg, nblocks * cfg.hop, seed=61) + workloads.noise_input(nblocks * cfg.hop, 62) * np.float32(0.1)
    chan = make_gpu_chain(FDC, cfg)
    halo, new = sharded.shard_input(x, cfg.hop, cfg.ovl, rank * per, per)
    chan.seek(rank * per, halo)
    sinks = sharded.ChannelSinks(chan, per, rank, world)
    d_in = torch.from_numpy(np.ascontiguousarray(new).view(np.float32).copy()).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    sinks.step(d_in.data_ptr(), per, stream)
    torch.cuda.synchronize(); dist.barrier()
    whole, _ = make_gpu_chain(FDC, cfg).work_host(x)
    got = np.empty(sinks.slab_blocks * sinks.per_sink[rank], dtype=np.complex64)
    if got.size:
        FDC._cabi.check(FDC._cabi.lib().fdc_memcpy_d2h(got.ctypes.data, sinks._local, got.nbytes))
    off = 0
    for i in sinks.my_channels:
        n = nblocks * cfg.params[i][2]
        assert np.array_equal(got[off:off + n].view(np.uint32), whole[i].view(np.uint32)), "rank %d channel %d" % (rank, i)
        off += n
    assert off == got.size and len(sinks.my_channels) >= 1
    dist.barrier(); sinks.close()

    # ---- 2. instance-sharded activity sinks ----
    N, R = 1024, 4
    calls = (24 * world, world, 31 * world + 1)
    total = sum(calls)
    spec, _ = sc.bursty_spectra(N, total, 8, seed=52, widths=(16, 32, 64), mean_on=9, mean_off=12)
    spec = np.ascontiguousarray(spec, dtype=np.complex64)
    # a keyed carrier inside the band the PowerActivationChannel instance watches, so that it really toggles
    spec[:, 400:520] *= np.where((np.arange(total) // 7) % 2, 8.0, 1.0).astype(np.float32)[:, None]

    def blocks():
        return [FDC.SegmentDetection(5, N, R, 0.1, 0.9, 10.0, 0.0312, 0.2, 4, 1, True, False, "", False, 0),
                FDC.SegmentDetection(6, N, R, 0.05, 0.5, 10.0, 0.0312, 0.2, 16, 0, True, False, "", False, 0),
                FDC.PowerActivationChannel(N, 0.45, 0.1, R, 6.0, 3, 1, True, False, "", 0, 7)]
    blks = blocks()
    owners = sharded.PeerBuffers(8 << 20, rank, world)
    grp = sharded.ShardedActivityGroup(blks, rank, world, owners=owners)
    mine = [[] for _ in blks]
    pos = 0
    d_all = torch.from_numpy(spec.view(np.float32).copy()).to(dev)
    for n in calls:
        first, count = sharded.partition(n, world)
        r0 = pos + first[rank]
        own = d_all.data_ptr() + 8 * N * r0
        prev = d_all.data_ptr() + 8 * N * (r0 - 1) if r0 > pos or pos > 0 else 0
        # the row before a rank's run: the previous row of the stream (the block's own history for the first rank of the first call)
        res = grp.work(n, own, prev if r0 > 0 else 0)
        for i, m in enumerate(res):
            assert (m is not None) == (i % world == rank)
            if m is not None:
                mine[i] += m
        pos += n
    torch.cuda.synchronize(); dist.barrier()
    ref_blks = blocks()
    for i, b in enumerate(ref_blks):
        if i % world != rank:
            continue
        b.work(total, [spec.reshape(-1)])
        want = b.messages()
        assert len(want) >= 2, (i, len(want), b.state() if hasattr(b, 'state') else None)
        assert [sc.meta_tuple(m) for m in mine[i]] == [sc.meta_tuple(m) for m in want], "instance %d metadata" % i
        for u, v in zip(mine[i], want):
            assert np.array_equal(np.asarray(u["data"]).view(np.uint32), np.asarray(v["data"]).view(np.uint32)), "instance %d samples" % i
    dist.barrier(); owners.close()
    if rank == 0:
        print("multirank ok: %d ranks" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
