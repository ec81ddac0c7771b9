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

import os
from dataclasses import dataclass, field
from typing import Any, Callable


TUNE_PEER_SCATTER_CTAS, TUNE_PEER_SCATTER_KERNEL = 6, 7  # enum b2_tunable (include/b200olap.h)


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
    all-to-all and the receiver skips its first partitioning pass (b2_join_pairs_seg_cap_phased_dev).

    One step enqueues, with NO host synchronisation and no eager tensor arithmetic in between:
      counts | all-gather of the boundaries (also the "everyone has read its receive buffers" barrier) |
      b2_shuffle_p2p_plan_dev per side (addresses, coarse boundaries, the collective overflow flag — all
      on the device) | build side over NVLink | barrier | on a second stream: the probe side over NVLink
      in `probe_shares` shares, a barrier and an event after each | on the main stream: the build side's
      fine pass (UNDER the first share's transfer), then per share: wait for its event, fine pass +
      probe (under the next share's transfer).
    The probe side's scatter kernel runs with a CTA budget while it shares the GPU (b2_tunable
    B2_TUNE_PEER_SCATTER_CTAS). Receive buffers are torch symmetric-memory allocations; their peer
    addresses come from the rendezvous handle. This class is plumbing only: every kernel is behind
    include/b200olap.h."""

    BITS = 10  # log2(G) destination bits + coarse bits: one radix pass

    def __init__(self, ctx, dist, rank: int, world: int, n_local: int, capacity: int, n_build_total: int | None = None,
                 probe_shares: int | None = None):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self.skip = log2_exact(world)
        self.seg_bits = self.BITS - self.skip
        self.capacity = capacity
        # build rows a rank receives when the hash spreads evenly: picks the fine partition count
        self.nr_expected = (n_build_total if n_build_total is not None else n_local * world) // world
        # shares of the probe side (sides 0 .. S-1; side S is the build side). One share is the measured
        # best: 8 B200, SF=2048, 96-CTA budget: 19.0 ms per step with one share, 19.8 ms with two (the
        # probe kernel builds its tables once per share and the concurrent kernels take SMs from each
        # other), 22.2 ms with three (profiles/r2_multi_gpu.md).
        S = probe_shares if probe_shares is not None else int(os.environ.get("B2_PROBE_SHARES", "1"))
        self.shares = S = max(1, S)
        self.share_rows = n_local if S == 1 else ((n_local + S - 1) // S + 1023) // 1024 * 1024
        cap_share = capacity if S == 1 else (capacity + S - 1) // S + 65536
        self.caps = [cap_share] * S + [capacity]
        dev = torch.device("cuda", torch.cuda.current_device())
        group = dist.group.WORLD.group_name
        B = 1 << self.BITS
        self.recv, self.peers, self.ws = [], [], []
        for side in range(S + 1):
            t = symm.empty(self.caps[side], dtype=torch.int64, device=dev)
            hdl = symm.rendezvous(t, group)
            self.recv.append(t)
            self.peers.append(torch.tensor(list(hdl.buffer_ptrs), dtype=torch.int64, device=dev))
            n_side = n_local if side == S else min(self.share_rows, n_local)
            self.ws.append(torch.empty(ctx.shuffle_p2p_ws_bytes(n_side, self.BITS) + 256, dtype=torch.uint8, device=dev))
        self.my_off = torch.zeros((S + 1, B + 1), dtype=torch.int64, device=dev)           # count kernels write here
        self.all_off = torch.zeros((world, S + 1, B + 1), dtype=torch.int64, device=dev)   # after the all-gather
        base = self.all_off.data_ptr()
        stride_rank, stride_side = (S + 1) * (B + 1) * 8, (B + 1) * 8
        self.off_ptrs = [torch.tensor([base + s * stride_rank + side * stride_side for s in range(world)],
                                      dtype=torch.int64, device=dev) for side in range(S + 1)]
        self.addr = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(S + 1)]
        self.seg = [torch.empty((1 << self.seg_bits) + 1, dtype=torch.int64, device=dev) for _ in range(S + 1)]
        self.info = torch.zeros((S + 1, 3), dtype=torch.int64, device=dev)  # {received, max over ranks, overflow} x side
        self.flag = torch.zeros(S + 1, dtype=torch.int32, device=dev)
        self.side = torch.cuda.Stream(device=dev)  # the probe side's scatters run here
        self.ev_r = torch.cuda.Event()
        self.ev_l = [torch.cuda.Event() for _ in range(S)]
        self.overlap = True
        # CTA budget of the probe side's scatter while the fine passes / probe want SMs too (0 = all).
        # Measured on 8 B200 at SF=2048 (tools/n8_overlap_sweep.sh, one share): all SMs 20.1 ms, 96 CTAs
        # 19.0 ms, 64 CTAs 22.2 ms per join step; with 2 ranks the scatter is not link-bound.
        self.probe_scatter_ctas = int(os.environ.get("B2_PROBE_SCATTER_CTAS", "96" if world >= 8 else "0"))
        if os.environ.get("B2_PEER_SCATTER_KERNEL"):  # 0 = whole lines stored by the threads, 1 = bulk sectors
            ctx.set_tunable(TUNE_PEER_SCATTER_KERNEL, int(os.environ["B2_PEER_SCATTER_KERNEL"]))
        self.last_recv = (0, 0)

    @property
    def abort(self):
        """int64[1] device view: non-zero = a receive buffer would have overflowed (on every rank)."""
        return self.info[self.shares, 2:3]

    def received(self) -> tuple[int, int]:
        """(L rows, R rows) this rank received in the last step — a host read, for reports only."""
        h = self.info.cpu()
        S = self.shares
        if int(h[S, 2]):
            raise OverflowError(f"a rank would receive {int(h[:, 1].max())} rows of one side / share, capacities "
                                f"{self.caps} (skewed keys); raised on every rank")
        self.last_recv = (int(h[:S, 0].sum()), int(h[S, 0]))
        return self.last_recv

    def _share(self, t, i: int):
        return t if self.shares == 1 else t[i * self.share_rows: min((i + 1) * self.share_rows, t.numel())]

    def step(self, fk, y, pk, x, local_join, phases: dict | None = None, probe_lt: int | None = None):
        """local_join(l_buf, l_seg_off, r_buf, r_seg_off, nr_expected, seg_bits, skip_bits, abort, phase_bits) ->
        result, e.g. ctx.join_pairs_seg_cap_dev(..., phases=phase_bits): 1 = build, 2 = probe one share,
        4 = finish. phases: if given, every phase of the step is synchronised and its wall time (ms) stored
        there (diagnostics only — the synchronisation removes all overlap). probe_lt: a predicate on the
        probe side's payload pushed in front of the link (filter -> join pipelines): rows with y >= probe_lt
        are neither counted nor sent (b2_shuffle_p2p_count_lt_dev / b2_shuffle_p2p_scatter_lt_dev)."""
        import time

        import torch
        ctx, G, S = self.ctx, self.world, self.shares
        t0 = [time.perf_counter()]

        def mark(name):
            if phases is not None:
                torch.cuda.synchronize()
                now = time.perf_counter()
                phases[name] = phases.get(name, 0.0) + (now - t0[0]) * 1e3
                t0[0] = now

        def join_phase(i: int, bits: int):
            return local_join(self.recv[i], self.seg[i], self.recv[S], self.seg[S], self.nr_expected, self.seg_bits,
                              self.skip, self.abort, bits)

        if phases is not None:
            torch.cuda.synchronize()
            t0[0] = time.perf_counter()
        for i in range(S):
            ctx.shuffle_p2p_count_dev(self._share(fk, i), self.BITS, self.ws[i], self.my_off[i],
                                      val=self._share(y, i), val_lt=probe_lt)
        ctx.shuffle_p2p_count_dev(pk, self.BITS, self.ws[S], self.my_off[S])
        mark("count")
        # also a barrier: nobody scatters before every rank is done reading its receive buffers
        self.dist.all_gather_into_tensor(self.all_off.view(-1), self.my_off.view(-1))
        for side in range(S + 1):  # the overflow flag chains through the sides: the last one covers all
            ctx.shuffle_p2p_plan_dev(self.off_ptrs[side], self.peers[side], self.rank, G, self.BITS, self.caps[side],
                                     self.addr[side], self.seg[side], self.info[side],
                                     prev_abort=None if side == 0 else self.info[side - 1, 2:3])
        mark("allgather+plan")
        if not self.overlap or phases is not None:
            ctx.shuffle_p2p_scatter_dev(pk, x, self.BITS, self.addr[S], self.ws[S], abort=self.abort)
            for i in range(S):
                ctx.shuffle_p2p_scatter_dev(self._share(fk, i), self._share(y, i), self.BITS, self.addr[i], self.ws[i],
                                            abort=self.abort, val_lt=probe_lt)
            mark("scatter_nvlink")
            self.dist.all_reduce(self.flag[0:1])  # every rank's stores have landed when this completes
            mark("barrier")
            out = None
            for i in range(S):
                out = join_phase(i, (1 if i == 0 else 0) | 2 | (4 if i == S - 1 else 0))
            mark("local_join")
            return out
        main = torch.cuda.current_stream()
        ctx.shuffle_p2p_scatter_dev(pk, x, self.BITS, self.addr[S], self.ws[S], abort=self.abort)
        self.ev_r.record(main)
        self.dist.all_reduce(self.flag[S:S + 1])       # build rows of every rank have landed
        self.side.wait_event(self.ev_r)                # the link is the build side's until then
        with torch.cuda.stream(self.side):
            for i in range(S):
                if self.probe_scatter_ctas:
                    ctx.set_tunable(TUNE_PEER_SCATTER_CTAS, self.probe_scatter_ctas)
                ctx.shuffle_p2p_scatter_dev(self._share(fk, i), self._share(y, i), self.BITS, self.addr[i], self.ws[i],
                                            abort=self.abort, val_lt=probe_lt)
                if self.probe_scatter_ctas:
                    ctx.set_tunable(TUNE_PEER_SCATTER_CTAS, 0)
                self.dist.all_reduce(self.flag[i:i + 1])   # this share's rows of every rank have landed
                self.ev_l[i].record(self.side)
        join_phase(0, 1)                               # build side's fine pass, under share 0's transfer
        out = None
        for i in range(S):
            main.wait_event(self.ev_l[i])
            out = join_phase(i, 2 | (4 if i == S - 1 else 0))
        return out
