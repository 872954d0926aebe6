"""Time sharding of the channelizer across the GPUs of one node (one process per GPU, torch.distributed).

The input stream is cut into contiguous runs of overlap-save blocks, one run per rank.  Block b depends only on the
samples [b*hop - ovl, b*hop + hop) (lib/overlap_save_impl.cc:70-78 in the reference) and the phase table index is the
closed form (b * shift) mod R (lib/phase_shifting_windowing_vcc_impl.cc:82), so a rank needs nothing from its
neighbours except its own ovl-sample halo, which it reads from the input itself: NO collective on the data path.
A collective is used only to bring the per-channel output runs (and, for the activity-gated blocks, the burst
metadata) to the rank that feeds the flowgraph sink: `gather_outputs` / `gather_objects`.

The per-rank compute is a callable so that this plumbing is testable on CPU (gloo, world_size 2) with the oracle
as the worker; on the GPU box the worker is `channelizer_worker(FDC.Channelizer(...))`.
"""
import numpy as np


def partition(nblocks, world):
    """Contiguous runs of blocks: rank r owns [first[r], first[r] + count[r]); sizes differ by at most one."""
    base, extra = divmod(int(nblocks), int(world))
    count = [base + (1 if r < extra else 0) for r in range(world)]
    first = [sum(count[:r]) for r in range(world)]
    return first, count


def shard_input(x, hop, ovl, first_block, nblocks, stream_history=None):
    """The samples rank needs: (halo[ovl], new[nblocks*hop]).  The halo of the very first block of the stream is the
    saved history of the previous call (zeros at stream start, lib/overlap_save_impl.cc:52)."""
    x = np.asarray(x)
    lo = first_block * hop
    new = x[lo:lo + nblocks * hop]
    if ovl == 0:
        return x[:0], new
    if lo >= ovl:
        return x[lo - ovl:lo], new
    hist = np.zeros(ovl, dtype=x.dtype) if stream_history is None else np.asarray(stream_history, dtype=x.dtype)
    assert hist.size == ovl
    return np.concatenate([hist[lo:], x[:lo]])[-ovl:] if lo else hist, new


def channelizer_worker(chan, global_block0=0):
    """worker(halo, new, first_block) -> list of per-channel arrays, on a FDC.Channelizer (CUDA)"""
    def work(halo, new, first_block):
        chan.seek(global_block0 + first_block, halo if chan.ovl else None)
        outs, _ = chan.work_host(new)
        return outs
    return work


def run_sharded(worker, x, hop, ovl, nblocks, rank, world, stream_history=None):
    """This rank's part of a call over `nblocks` blocks of the stream x (every rank holds or can read x)."""
    first, count = partition(nblocks, world)
    halo, new = shard_input(x, hop, ovl, first[rank], count[rank], stream_history)
    return worker(halo, new, first[rank]) if count[rank] else None


def gather_outputs(local_outs, louts, nblocks, rank, world, dst=0, group=None, device=None):
    """Concatenate every rank's per-channel runs on rank `dst` in stream order.  Run lengths follow from `partition`, so
    no size exchange is needed; one gather per call (equal-sized slabs are padded by at most one block)."""
    import torch
    import torch.distributed as dist
    first, count = partition(nblocks, world)
    per_block = int(sum(louts))
    maxb = max(count)
    dev = device if device is not None else torch.device("cpu")
    slab = torch.zeros(maxb * per_block * 2, dtype=torch.float32, device=dev)
    if local_outs is not None and count[rank]:
        flat = np.concatenate([np.ascontiguousarray(o, dtype=np.complex64) for o in local_outs]).view(np.float32)
        slab[:flat.size] = torch.from_numpy(flat).to(dev)
    bufs = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
    dist.gather(slab, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    outs = [[] for _ in louts]
    for r in range(world):
        if not count[r]:
            continue
        a = bufs[r].cpu().numpy().view(np.complex64)
        off = 0
        for i, lo in enumerate(louts):
            outs[i].append(a[off:off + count[r] * lo]); off += count[r] * lo
    return [np.concatenate(o) if o else np.zeros(0, dtype=np.complex64) for o in outs]


def gather_objects(obj, rank, world, dst=0, group=None):
    """Variable-length metadata (burst PDUs of the activity-gated blocks): gathered as Python objects in rank order."""
    import torch.distributed as dist
    res = [None] * world if rank == dst else None
    dist.gather_object(obj, res, dst=dst, group=group)
    return res


class PeerSink(object):
    """One stream-ordered output buffer on the sink rank that every rank's extract kernel stores into directly (peer memory
    over NVLink, CUDA IPC between the per-GPU processes): the gather is fused into the kernel, nothing is staged locally and
    no collective runs afterwards.  Layout: channel-major slabs of world * blocks_per_rank rows; rank r owns rows
    [r * blocks_per_rank, (r + 1) * blocks_per_rank) of every slab (Channelizer.work_device_slab)."""

    def __init__(self, out_per_block, blocks_per_rank, rank, world, dst=0, group=None):
        import ctypes as C
        import torch.distributed as dist
        from ._cabi import lib, check, handle
        self.rank, self.world, self.dst, self.blocks_per_rank = rank, world, dst, int(blocks_per_rank)
        self.slab_blocks = world * self.blocks_per_rank
        self.nbytes = 8 * self.slab_blocks * int(out_per_block)
        self._local = None
        box = [None]
        if rank == dst:
            self._local = lib().fdc_dev_alloc(self.nbytes)
            if not self._local:
                raise RuntimeError("sink allocation failed")
            h = C.create_string_buffer(64)
            check(lib().fdc_ipc_export(C.c_void_p(self._local), h), "fdc_ipc_export")
            box[0] = bytes(h.raw)
        dist.broadcast_object_list(box, src=dst, group=group)
        if rank == dst:
            self.ptr = self._local
        else:
            self._hbuf = C.create_string_buffer(box[0], 64)
            self.ptr = handle(lib().fdc_ipc_open(self._hbuf), "fdc_ipc_open").value

    def first_block(self):
        return self.rank * self.blocks_per_rank

    def close(self):
        from ._cabi import lib
        if self.ptr and self.rank != self.dst:
            lib().fdc_ipc_close(self.ptr)
        if self._local:
            lib().fdc_dev_free(self._local)
        self.ptr = self._local = None
