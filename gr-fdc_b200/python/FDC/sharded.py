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

    def __init__(self, out_per_block, blocks_per_rank, rank, world, dst=0, group=None, nbytes=None):
        """nbytes: a plain buffer of that size instead of the channel-major slab layout (ShardedActivityGroup)."""
        import ctypes as C
        import torch.distributed as dist
        from ._cabi import lib, check, handle
        self.rank, self.world, self.dst, self.blocks_per_rank = rank, world, dst, int(blocks_per_rank)
        self.slab_blocks = world * self.blocks_per_rank
        self.nbytes = int(nbytes) if nbytes is not None else 8 * self.slab_blocks * int(out_per_block)
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


def channel_owners(louts, world):
    """Deal the channels out over `world` sinks in contiguous runs of (nearly) equal output volume: owner[i] for every channel and
    the number of output items per block each sink receives.  Contiguous runs keep neighbouring channels (which share spectrum
    bins and usually a downstream consumer) on one GPU."""
    louts = [int(v) for v in louts]
    total = sum(louts)
    owner, per_sink, acc, k = [], [0] * world, 0, 0
    for lo in louts:
        # move on to the next sink when this one has its share (never leave a later sink without channels to take)
        while k < world - 1 and acc + lo / 2.0 > total * (k + 1) / float(world):
            k += 1
        owner.append(k); per_sink[k] += lo; acc += lo
    return owner, per_sink


class ChannelSinks(object):
    """Channel-sharded sinks of a time-sharded channelizer (one process per GPU).  Rank k owns the channels with owner[i] == k and
    holds ONE buffer for them: channel-major slabs of world * blocks_per_rank rows.  Every rank maps every other rank's buffer
    (CUDA IPC, peer access over NVLink) and its extract kernel stores each channel's rows [rank * blocks_per_rank, ...) straight
    into the owner's buffer: an all-to-all of peer stores fused into the kernel; each GPU receives (world-1)/world of ONE GPU's
    output instead of a single sink rank receiving (world-1) GPUs' worth (PeerSink, NCCL gather).  After a step and a barrier,
    rank k holds the complete, stream-ordered output of its channels."""

    def __init__(self, chan, blocks_per_rank, rank, world, group=None):
        import ctypes as C
        import torch.distributed as dist
        from ._cabi import lib, check, handle
        self.rank, self.world, self.blocks_per_rank = rank, world, int(blocks_per_rank)
        self.slab_blocks = world * self.blocks_per_rank
        self.owner, self.per_sink = channel_owners(chan.lout, world)
        self.my_channels = [i for i, o in enumerate(self.owner) if o == rank]
        self.nbytes = 8 * self.slab_blocks * max(1, self.per_sink[rank])
        self._local = lib().fdc_dev_alloc(self.nbytes)
        if not self._local:
            raise RuntimeError("sink allocation failed")
        h = C.create_string_buffer(64)
        check(lib().fdc_ipc_export(C.c_void_p(self._local), h), "fdc_ipc_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        self._hbufs, self.bases = [], []
        for r in range(world):
            if r == rank:
                self.bases.append(self._local)
            else:
                hb = C.create_string_buffer(handles[r], 64); self._hbufs.append(hb)
                self.bases.append(handle(lib().fdc_ipc_open(hb), "fdc_ipc_open").value)
        chan.set_sinks(self.bases, self.owner, rank)
        self.chan = chan

    def first_block(self):
        return self.rank * self.blocks_per_rank

    def step(self, d_in, nblocks, stream=0):
        self.chan.work_device_sinks(d_in, nblocks, self.slab_blocks, self.first_block(), stream)

    def close(self):
        from ._cabi import lib
        for r, b in enumerate(self.bases):
            if r != self.rank and b:
                lib().fdc_ipc_close(b)
        if self._local:
            lib().fdc_dev_free(self._local)
        self.bases, self._local = [], None


class PeerBuffers(object):
    """One plain device buffer per rank, every rank mapping all of them (CUDA IPC, peer access over NVLink): ptrs[r] is rank r's
    buffer as seen from this GPU.  Used by ShardedActivityGroup to deal the activity-gated block instances out over the ranks:
    rank r receives the burst samples of the instances it owns and assembles their PDUs."""

    def __init__(self, nbytes, rank, world, group=None):
        import ctypes as C
        import torch.distributed as dist
        from ._cabi import lib, check, handle
        self.rank, self.world, self.nbytes = rank, world, int(nbytes)
        self._local = lib().fdc_dev_alloc(self.nbytes)
        if not self._local:
            raise RuntimeError("peer buffer allocation failed")
        h = C.create_string_buffer(64)
        check(lib().fdc_ipc_export(C.c_void_p(self._local), h), "fdc_ipc_export")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        self._hbufs, self.ptrs = [], []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self._local)
            else:
                hb = C.create_string_buffer(handles[r], 64); self._hbufs.append(hb)
                self.ptrs.append(handle(lib().fdc_ipc_open(hb), "fdc_ipc_open").value)

    def close(self):
        from ._cabi import lib
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                lib().fdc_ipc_close(p)
        if self._local:
            lib().fdc_dev_free(self._local)
        self.ptrs, self._local = [], None


class ShardedActivity(object):
    """An activity-gated block (PowerActivationChannel, SegmentDetection, activity_detection_channelizer_vcm) on a
    time-sharded stream.  The reference blocks carry state from block to block (lib/PowerActivationChannel_impl.cc:137-177,
    lib/SegmentDetection_impl.cc:163-300); their per-block measurements are stateless.  So every rank measures its own rows
    on its GPU, the compact records (a few bytes per block) are all-gathered, EVERY rank runs the same sequential bookkeeping
    over them and thereby knows the whole job list, each rank extracts the jobs its rows emitted, and only those samples
    travel to the sink rank, which publishes the PDUs in the reference's order.

    Every rank builds `block` with the same constructor arguments.  Rows: the fft-shifted, 1/N-scaled spectrum blocks the
    reference block would see on its input."""

    def __init__(self, block, rank, world, dst=0, group=None):
        self.block, self.rank, self.world, self.dst, self.group = block, int(rank), int(world), int(dst), group

    def work(self, nblocks_total, d_rows=0, d_prev=0, stream=0, power=None):
        """One global call over nblocks_total blocks; this rank holds rows partition(nblocks_total, world)[rank] at d_rows
        and the row before its first one at d_prev (0: all-zero row, start of the stream).  Host-logic contexts pass this
        rank's `power` rows instead.  Returns the messages on the sink rank, None elsewhere."""
        g = ShardedActivityGroup([self.block], self.rank, self.world, self.dst, self.group)
        res = g.work(nblocks_total, d_rows, d_prev, stream, powers=None if power is None else [power])
        return None if res is None else res[0]


class ShardedActivityGroup(object):
    """Several activity-gated blocks fed by the same spectrum (the hier block hangs 2 SegmentDetection and 16
    PowerActivationChannel blocks on one FFT in configs[2]): ONE all-gather of all blocks' records and ONE gather of all
    blocks' samples per call.  `pool` (a concurrent.futures executor) runs the per-block local phases concurrently, as
    GNU Radio's thread-per-block scheduler would; the collectives are issued from the calling thread only."""

    def __init__(self, blocks, rank, world, dst=0, group=None, pool=None, sink=None, arrays=False, owners=None):
        """sink: a PeerSink(nbytes=...) made by all ranks -- the extract kernels then store the burst samples straight into
        the sink rank's buffer (peer memory over NVLink) and the gather of the samples is a barrier; calls whose samples do
        not fit the buffer fall back to the NCCL gather.
        owners: a PeerBuffers made by all ranks -- block instance i is then OWNED by rank i % world: the extract kernels of all
        ranks store instance i's burst samples into its owner's buffer and the owner assembles and publishes its PDUs (the PDU
        assembly, serial on one sink rank otherwise, is spread over the ranks).  work() returns the messages of the owned
        instances and None for the others."""
        self.blocks, self.rank, self.world, self.dst, self.group = list(blocks), int(rank), int(world), int(dst), group
        self.sink = sink
        self.owners = owners
        self.arrays = bool(arrays)          # work() returns messages_arrays(reuse=True) per block instead of lists of dicts
        self._map = pool.map if pool is not None else (lambda f, it: list(map(f, it)))
        self.phase_seconds = {}                 # wall clock per phase, accumulated over calls (for measurements)

    def work(self, nblocks_total, d_rows=0, d_prev=0, stream=0, powers=None):
        import time
        import torch.distributed as dist
        first, count = partition(nblocks_total, self.world)
        me, nb = self.rank, len(self.blocks)
        idx = range(nb)
        t = [time.perf_counter()]
        recs = list(self._map(lambda i: self.blocks[i].shard_measure(count[me], d_rows, stream, power=None if powers is None else powers[i]), idx))
        t.append(time.perf_counter())
        allrecs = _all_gather_bytes(recs, self.world, self.group)
        t.append(time.perf_counter())

        def decide(i):
            b = self.blocks[i]
            b.shard_decide(nblocks_total, b"".join(allrecs[r][i] for r in range(self.world)))
            return [b.shard_samples(first[r], count[r]) for r in range(self.world)]
        sizes = list(self._map(decide, idx))                                        # [block][rank], known everywhere: no size exchange
        total = [sum(sz) for sz in sizes]
        t.append(time.perf_counter())
        out = None
        own_of = [i % self.world for i in idx]
        per_owner = [sum(total[i] for i in idx if own_of[i] == r) for r in range(self.world)]
        if self.owners is not None and 8 * max(per_owner + [0]) <= self.owners.nbytes:
            # instance-major layout in each owner's buffer; inside an instance's run the samples of ALL ranks lie channel by channel
            # (shard_layout: every rank derives the same offsets), so that the owner can publish whole bursts as views
            base, run = [], [0] * self.world
            for i in idx:
                base.append(run[own_of[i]]); run[own_of[i]] += total[i]

            def extract_to_owner(i):
                self.blocks[i].shard_layout(True)
                return self.blocks[i].shard_extract_device(first[me], count[me], d_rows, d_prev, self.owners.ptrs[own_of[i]] + 8 * base[i], stream)
            list(self._map(extract_to_owner, idx))
            t.append(time.perf_counter())
            dist.barrier(group=self.group)                                          # every rank's stores have landed
            t.append(time.perf_counter())

            def assemble_own(i):
                b = self.blocks[i]
                if own_of[i] != me:
                    b.shard_assemble(None)
                    return None
                b.shard_assemble_device(self.owners.ptrs[me] + 8 * base[i], total[i], stream)
                return b.messages_arrays(reuse=True) if self.arrays else b.messages()
            out = list(self._map(assemble_own, idx))
        elif self.sink is not None and 8 * sum(total) <= self.sink.nbytes:
            # instance-major layout in the sink's buffer, all ranks' samples of an instance channel by channel (shard_layout)
            base = np.concatenate([[0], np.cumsum(total)]).tolist()

            def extract_to_sink(i):
                self.blocks[i].shard_layout(True)
                return self.blocks[i].shard_extract_device(first[me], count[me], d_rows, d_prev, self.sink.ptr + 8 * base[i], stream)
            list(self._map(extract_to_sink, idx))
            t.append(time.perf_counter())
            dist.barrier(group=self.group)                                          # every rank's stores have landed
            t.append(time.perf_counter())
            if me == self.dst:
                def assemble_dev(i):
                    self.blocks[i].shard_assemble_device(self.sink.ptr + 8 * base[i], total[i], stream)
                    return self.blocks[i].messages_arrays(reuse=True) if self.arrays else self.blocks[i].messages()
                out = list(self._map(assemble_dev, idx))
            else:
                for b in self.blocks:
                    b.shard_assemble(None)
        else:
            res = list(self._map(lambda i: self.blocks[i].shard_extract(first[me], count[me], d_rows, d_prev, stream), idx))
            t.append(time.perf_counter())
            mine = np.concatenate(res) if nb else np.zeros(0, np.complex64)
            parts = _gather_ragged(mine, [sum(sizes[i][r] for i in idx) for r in range(self.world)], me, self.world, self.dst, self.group)
            t.append(time.perf_counter())
            if me != self.dst:
                for b in self.blocks:
                    b.shard_assemble(None)
            else:
                cuts = [np.concatenate([[0], np.cumsum([sizes[i][r] for i in idx])]).tolist() for r in range(self.world)]

                def assemble(i):
                    b = self.blocks[i]
                    b.shard_assemble(np.concatenate([parts[r][cuts[r][i]:cuts[r][i + 1]] for r in range(self.world)]))
                    return b.messages_arrays(reuse=True) if self.arrays else b.messages()
                out = list(self._map(assemble, idx))
        t.append(time.perf_counter())
        for k, name in enumerate(("measure", "allgather_records", "decide", "extract", "gather_samples", "assemble")):
            self.phase_seconds[name] = self.phase_seconds.get(name, 0.0) + t[k + 1] - t[k]
        return out


def _all_gather_bytes(recs, world, group):
    """All-gather of every rank's list of byte strings (the detection records): sizes first, then one padded all_gather --
    two small collectives instead of a pickled object gather."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    lens = torch.tensor([len(r) for r in recs], dtype=torch.int64, device=dev)
    all_lens = [torch.empty_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    all_lens = [a.cpu().numpy() for a in all_lens]
    width = max(int(max(a.sum() for a in all_lens)), 1)
    buf = torch.zeros(width, dtype=torch.uint8)
    flat = b"".join(recs)
    if flat:
        buf[:len(flat)] = torch.frombuffer(bytearray(flat), dtype=torch.uint8)
    buf = buf.to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    res = []
    for r in range(world):
        raw = outs[r].cpu().numpy().tobytes()
        cut = np.concatenate([[0], np.cumsum(all_lens[r])])
        res.append([raw[cut[i]:cut[i + 1]] for i in range(len(recs))])
    return res


def _gather_ragged(mine, sizes, rank, world, dst, group):
    """Gather complex64 runs of known sizes on rank dst (one padded gather; the NCCL path when the group is NCCL)."""
    import torch
    import torch.distributed as dist
    width = max(max(sizes), 1)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    slab = torch.zeros(2 * width, dtype=torch.float32)
    if mine.size:
        slab[:2 * mine.size] = torch.from_numpy(np.ascontiguousarray(mine).view(np.float32))
    slab = slab.to(dev)
    bufs = [torch.empty_like(slab) for _ in range(world)] if rank == dst else None
    dist.gather(slab, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [bufs[r].cpu().numpy().view(np.complex64)[:sizes[r]].copy() for r in range(world)]
