"""Multi-GPU orchestration of the operator path (one process per GPU, torch.distributed).

What shards how (the reference's only parallelism is "batch i -> DPU i",
filter_dpu.cc:127, and a host-mediated repartition for the join, partitioner.cc:350-375):

* filter / sum / take: contiguous batch ranges per rank, no data-path collective
  (:func:`shard_range`); the sum adds one partial per rank (:func:`all_sum_u64`).
* join: both sides are routed by the top log2(G) bits of wang_hash(key), the (key, payload) pairs
  cross NVLink in ONE exchange per side, every rank joins what it received
  (:class:`ShardedJoin`).

The device work (routing kernel, local join) is injected, so the exchange plumbing — counts,
split sizes, capacity checks, the all-to-all itself — runs unchanged on CPU tensors over gloo
(tests/test_sharded_gloo.py) and on CUDA tensors over NCCL (bench.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable


def shard_range(nbatches: int, rank: int, world: int) -> tuple[int, int]:
    """(first batch, batch count) of `rank`: contiguous ranges, remainder to the low ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    per, rem = divmod(nbatches, world)
    first = rank * per + min(rank, rem)
    return first, per + (1 if rank < rem else 0)


def log2_exact(n: int) -> int:
    if n < 1 or n & (n - 1):
        raise ValueError(f"the sharded join needs a power-of-two number of ranks, got {n}")
    return n.bit_length() - 1


def all_sum_u64(dist: Any, partial: int, device: Any = "cpu") -> int:
    """Sum of uint64 partials over all ranks, mod 2^64 (SumDpu adds per-DPU partials on the host,
    aggr_dpu.cc:82-84). int64 two's-complement addition wraps exactly like uint64."""
    import torch
    if dist is None:
        return partial & 0xFFFFFFFFFFFFFFFF
    v = partial & 0xFFFFFFFFFFFFFFFF
    t = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item()) & 0xFFFFFFFFFFFFFFFF


@dataclass
class ExchangeCounts:
    send_l: list[int]
    send_r: list[int]
    recv_l: list[int]
    recv_r: list[int]

    @property
    def nl(self) -> int:
        return sum(self.recv_l)

    @property
    def nr(self) -> int:
        return sum(self.recv_r)


@dataclass
class ShardedJoin:
    """One join step of rank `rank` of `world`.

    route(key, val) -> (pairs int64[n], dest_off int64[G+1]): rows grouped by destination rank
        (b2_shuffle_partition_u32_dev on the GPU);
    local_join(l_pairs, r_pairs, skip_bits) -> result: joins the received pairs, ignoring the
        top skip_bits hash bits (b2_join_pairs_dev on the GPU).
    """
    dist: Any
    rank: int
    world: int
    route: Callable
    local_join: Callable
    capacity: int | None = None          # rows each receive buffer can hold (None: allocate exactly)
    route_r: Callable | None = None      # separate routing step for the build side (own buffers)
    recv_l: Any = None                   # optional preallocated int64 receive buffers
    recv_r: Any = None
    last: ExchangeCounts | None = field(default=None, init=False)

    def exchange_counts(self, off) -> tuple[list[int], list[int]]:
        """(send, recv) row counts per peer for one side: one small all-gather of every rank's send
        counts plus the host read that sizes the pair exchange (the reference reads the per-DPU
        histograms back to the host for the same purpose, partitioner.cc:167-180,280-312). Every
        rank sees the whole G x G matrix, so the capacity check below is decided identically on all
        ranks: either everybody raises or nobody does (a rank that raised alone would leave its
        peers hanging in the all-to-all)."""
        import torch
        counts = (off[1:] - off[:-1]).contiguous()
        if self.dist is None:
            allc = counts.view(1, -1)
        else:
            flat = torch.empty(self.world * counts.numel(), dtype=counts.dtype, device=counts.device)
            self.dist.all_gather_into_tensor(flat, counts)  # flat output: gloo wants it that way
            allc = flat.view(self.world, counts.numel())
        m = allc.cpu()
        worst = int(m.sum(0).max())
        self._check_capacity(worst)
        return m[self.rank].tolist(), m[:, self.rank].tolist()

    def _check_capacity(self, worst: int) -> None:
        caps = [c for c in (self.capacity,
                            None if self.recv_l is None else self.recv_l.numel(),
                            None if self.recv_r is None else self.recv_r.numel()) if c is not None]
        if caps and worst > min(caps):
            raise OverflowError(f"a rank would receive {worst} rows, capacity {min(caps)} (skewed keys); "
                                "raised on every rank")

    def _recv_buffer(self, pre, n, like):
        import torch
        if pre is not None:
            return pre[:n]  # capacity was checked collectively in exchange_counts
        return torch.empty(n, dtype=like.dtype, device=like.device)

    def exchange(self, pairs, send, recv, pre=None):
        """All-to-all of 8-byte pairs with the given per-peer split sizes, started asynchronously:
        returns (receive buffer, work handle or None). Work enqueued on the caller's stream after
        this call overlaps with the transfer until handle.wait()."""
        out = self._recv_buffer(pre, sum(recv), pairs)
        if self.dist is None:
            out.copy_(pairs[: sum(send)])
            return out, None
        work = self.dist.all_to_all_single(out, pairs[: sum(send)], output_split_sizes=recv,
                                           input_split_sizes=send, async_op=True)
        return out, work

    def step(self, fk, y, pk, x):
        """route L | exchange L (async) overlapped with route R | exchange R | local join.
        Returns local_join's result."""
        skip = log2_exact(self.world)
        lp, l_off = self.route(fk, y)
        send_l, recv_l = self.exchange_counts(l_off)
        lrecv, wl = self.exchange(lp, send_l, recv_l, self.recv_l)
        rp, r_off = (self.route_r or self.route)(pk, x)   # runs while the L pairs cross NVLink
        send_r, recv_r = self.exchange_counts(r_off)
        rrecv, wr = self.exchange(rp, send_r, recv_r, self.recv_r)
        self.last = ExchangeCounts(send_l, send_r, recv_l, recv_r)
        for w in (wl, wr):
            if w is not None:
                w.wait()
        return self.local_join(lrecv, rrecv, skip)

    def bytes_sent(self) -> int:
        """Bytes this rank put on the wire in the last step (pairs to OTHER ranks)."""
        c = self.last
        if c is None:
            return 0
        return 8 * (sum(c.send_l) - c.send_l[self.rank] + sum(c.send_r) - c.send_r[self.rank])


# -------------------------------------------------------------------------------------------------
# Fused shuffle over peer memory
# -------------------------------------------------------------------------------------------------
def p2p_plan(all_counts, peer_ptrs, rank: int, world: int):
    """Destination addresses of one side of the fused shuffle.

    all_counts int64[G, B]: rows of source rank s in bucket b (B = G * C buckets: destination rank
    b // C, coarse bucket b % C there); peer_ptrs int64[G]: base address of every rank's receive
    buffer. The receive buffer of rank r is laid out bucket-major, source-minor, so every coarse
    bucket is contiguous. Returns
      addr    int64[B]   byte address where THIS rank's rows of bucket b go,
      seg_off int64[C+1] boundaries (rows) of the coarse buckets THIS rank receives,
      n_recv  int64[]    rows THIS rank receives,
      max_recv int64[]   largest receive count over all ranks (for the capacity check).
    Pure tensor arithmetic: runs on the device without a host round trip, and on CPU tensors in
    the tests."""
    import torch
    G = world
    B = all_counts.shape[1]
    C = B // G
    tot = all_counts.sum(0).view(G, C)                        # rows per (destination, coarse bucket)
    base_in_dest = (tot.cumsum(1) - tot).reshape(B)           # rows before bucket b at its destination
    src_off = (all_counts.cumsum(0) - all_counts)[rank]       # rows of lower source ranks in bucket b
    addr = peer_ptrs.repeat_interleave(C) + 8 * (base_in_dest + src_off)
    seg_off = torch.zeros(C + 1, dtype=torch.int64, device=all_counts.device)
    seg_off[1:] = tot[rank].cumsum(0)
    per_rank = tot.sum(1)
    return addr, seg_off, per_rank[rank], per_rank.max()


class P2PShuffleJoin:
    """Sharded join whose exchange is FUSED into the routing kernel: every rank scatters its
    (key, payload) pairs straight into the peers' receive buffers with NVLink stores
    (b2_shuffle_p2p_scatter_dev), already grouped into coarse partitions, so there is no separate
    all-to-all and the receiver skips its first partitioning pass (b2_join_pairs_seg_dev).
    Receive buffers are torch symmetric-memory allocations; their peer addresses come from the
    rendezvous handle. Collectives left: one all-gather of 2 x 1024 counts and one barrier."""

    BITS = 10  # log2(G) destination bits + coarse bits: one radix pass

    def __init__(self, ctx, dist, rank: int, world: int, n_local: int, capacity: int):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self.skip = log2_exact(world)
        self.seg_bits = self.BITS - self.skip
        self.capacity = capacity
        dev = torch.device("cuda", torch.cuda.current_device())
        group = dist.group.WORLD.group_name
        self.recv, self.peers = [], []
        for _ in range(2):  # L, R
            t = symm.empty(capacity, dtype=torch.int64, device=dev)
            hdl = symm.rendezvous(t, group)
            self.recv.append(t)
            self.peers.append(torch.tensor(list(hdl.buffer_ptrs), dtype=torch.int64, device=dev))
        nbytes = ctx.shuffle_p2p_ws_bytes(n_local, self.BITS) + 256
        self.ws = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.off = [torch.empty((1 << self.BITS) + 1, dtype=torch.int64, device=dev) for _ in range(2)]
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.last_recv = (0, 0)

    def step(self, fk, y, pk, x, local_join, phases: dict | None = None):
        """local_join(l_pairs, l_seg_off, r_pairs, r_seg_off, seg_bits, skip_bits) -> result.
        phases: if given, every phase is synchronised and its wall time (ms) stored there
        (diagnostics only — the synchronisation removes all overlap)."""
        import time

        import torch
        ctx, G = self.ctx, self.world
        t0 = [time.perf_counter()]

        def mark(name):
            if phases is not None:
                torch.cuda.synchronize()
                now = time.perf_counter()
                phases[name] = phases.get(name, 0.0) + (now - t0[0]) * 1e3
                t0[0] = now

        if phases is not None:
            torch.cuda.synchronize()
            t0[0] = time.perf_counter()
        ctx.shuffle_p2p_count_dev(fk, self.BITS, self.ws[0], self.off[0])
        ctx.shuffle_p2p_count_dev(pk, self.BITS, self.ws[1], self.off[1])
        mark("count")
        counts = torch.stack([self.off[0][1:] - self.off[0][:-1], self.off[1][1:] - self.off[1][:-1]])
        allc = torch.empty((G,) + tuple(counts.shape), dtype=torch.int64, device=counts.device)
        # also a barrier: nobody scatters before every rank is done reading its receive buffers
        self.dist.all_gather_into_tensor(allc, counts)
        plans = [p2p_plan(allc[:, side, :].contiguous(), self.peers[side], self.rank, G) for side in range(2)]
        sizes = torch.stack([plans[0][2], plans[1][2], plans[0][3], plans[1][3]]).cpu().tolist()
        if max(sizes[2], sizes[3]) > self.capacity:
            raise OverflowError(f"rank {self.rank}: a rank would receive {max(sizes[2], sizes[3])} rows, "
                                f"capacity {self.capacity} (skewed keys)")
        mark("allgather+plan")
        ctx.shuffle_p2p_scatter_dev(fk, y, self.BITS, plans[0][0], self.ws[0])
        ctx.shuffle_p2p_scatter_dev(pk, x, self.BITS, plans[1][0], self.ws[1])
        mark("scatter_nvlink")
        self.dist.all_reduce(self.flag)  # every rank's stores have landed when this completes
        mark("barrier")
        nl, nr = int(sizes[0]), int(sizes[1])
        self.last_recv = (nl, nr)
        out = local_join(self.recv[0][:nl], plans[0][1], self.recv[1][:nr], plans[1][1], self.seg_bits,
                         self.skip)
        mark("local_join")
        return out
