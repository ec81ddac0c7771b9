"""World-size-2 tests of the multi-GPU orchestration (dpu_olap_b200/sharded.py) over gloo on CPU.

The exchange plumbing (counts all-to-all, split sizes, pair all-to-all, capacity checks, uint64
all-reduce) is the code bench.py runs over NCCL; here the two injected device steps (routing rows
by destination rank, joining the received pairs) are stood in for by the CPU oracle, which is
allowed in tests only. Routing uses the library's own host-callable b2_join_dest_rank for a sample
so the two agree on who owns a key.
"""
import os
import socket
import traceback

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

WORLD = 2


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = globals()[fn_name](rank, world, dist)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", out))
    except Exception:  # pragma: no cover - reported to the parent
        q.put((rank, "err", traceback.format_exc()))


def _run(fn_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, fn_name, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, status, out in res:
        assert status == "ok", f"rank {rank}:\n{out}"
    return {rank: out for rank, _, out in res}


# ---- stand-ins for the device steps (CPU oracle; tests only) ---------------------------------------
def _pack(keys, vals):
    return torch.from_numpy((keys.astype(np.uint64) | (vals.astype(np.uint64) << np.uint64(32))).view(np.int64))


def _unpack(pairs):
    a = pairs.numpy().view(np.uint64)
    return (a & np.uint64(0xFFFFFFFF)).astype(np.uint32), (a >> np.uint64(32)).astype(np.uint32)


def _make_route(world):
    import oracle

    def route(key, val):
        dest = oracle.partition_ids(key, world)  # top log2(world) bits of wang_hash, as the GPU kernel
        order = np.argsort(dest, kind="stable")
        off = np.zeros(world + 1, dtype=np.int64)
        off[1:] = np.cumsum(np.bincount(dest, minlength=world))
        return _pack(key[order], val[order]), torch.from_numpy(off)
    return route


def _local_join(lrecv, rrecv, skip_bits):
    import oracle
    fk, y = _unpack(lrecv)
    pk, x = _unpack(rrecv)
    return oracle.join(fk, y, pk, x), (fk, pk)


def _inputs(rank, world, n=20000):
    import oracle
    g = oracle.RandomArrayGenerator(42)
    nb, bs = 4, n // 4
    x = oracle.make_random_batches(g, nb, bs)
    y = oracle.make_random_batches(g, nb, bs)
    fk = oracle.make_fk_batches(g, bs, nb, bs)
    pk = oracle.make_index_batches(nb, bs)
    cols = [np.concatenate(c) for c in (fk, y, pk, x)]
    from dpu_olap_b200.sharded import shard_range
    first, cnt = shard_range(nb, rank, world)
    sl = slice(first * bs, (first + cnt) * bs)
    return cols, [c[sl] for c in cols]


# ---- bodies executed on every rank -----------------------------------------------------------------
def body_join(rank, world, dist):
    import oracle
    from dpu_olap_b200 import _lib
    from dpu_olap_b200.sharded import ShardedJoin
    full, (fk, y, pk, x) = _inputs(rank, world)
    sj = ShardedJoin(dist, rank, world, _make_route(world), _local_join)
    (o_fk, o_y, o_x), (got_fk, got_pk) = sj.step(fk, y, pk, x)
    # every received key belongs to this rank according to the library's own routing function
    h = _lib.lib()
    for k in list(got_fk[:200]) + list(got_pk[:200]):
        assert h.b2_join_dest_rank(int(k), world) == rank
    c = sj.last
    assert sum(c.send_l) == fk.size and sum(c.send_r) == pk.size
    assert sj.bytes_sent() == 8 * (fk.size - c.send_l[rank] + pk.size - c.send_r[rank])
    gathered = [None] * world
    dist.all_gather_object(gathered, (o_fk, o_y, o_x))
    got = oracle.sort_rows(*[np.concatenate([g[i] for g in gathered]) for i in range(3)])
    exp = oracle.sort_rows(*oracle.join(*full))
    assert all(np.array_equal(a, b) for a, b in zip(got, exp))
    assert got[0].size == full[0].size  # every fk matches exactly one pk (join_test.cc:115-116)
    return int(o_fk.size)


def body_overflow(rank, world, dist):
    from dpu_olap_b200.sharded import ShardedJoin
    _, (fk, y, pk, x) = _inputs(rank, world)
    sj = ShardedJoin(dist, rank, world, _make_route(world), _local_join, capacity=10)
    try:
        sj.step(fk, y, pk, x)
    except OverflowError as e:
        return str(e)
    raise AssertionError("a receive capacity of 10 rows must overflow")


def body_sum(rank, world, dist):
    from dpu_olap_b200.sharded import all_sum_u64
    partial = (1 << 63) + 5 + rank  # forces a wrap past 2^64
    return all_sum_u64(dist, partial)


# ---- tests -----------------------------------------------------------------------------------------
def test_shard_range_covers_everything():
    from dpu_olap_b200.sharded import shard_range
    for nb in (0, 1, 7, 8, 2048, 262144):
        for world in (1, 2, 4, 8):
            spans = [shard_range(nb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == nb
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_log2_exact_rejects_non_powers_of_two():
    from dpu_olap_b200.sharded import log2_exact
    assert [log2_exact(n) for n in (1, 2, 4, 8)] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        log2_exact(6)


def test_sharded_join_two_ranks_matches_oracle():
    out = _run("body_join")
    assert sum(out.values()) == 20000


def test_sharded_join_reports_receive_overflow():
    out = _run("body_overflow")
    assert all("skewed" in v for v in out.values())


def test_all_sum_u64_wraps_like_uint64():
    out = _run("body_sum")
    exp = (((1 << 63) + 5) + ((1 << 63) + 6)) & 0xFFFFFFFFFFFFFFFF
    assert out[0] == out[1] == exp


def test_p2p_plan_layout_is_a_partition_of_every_receive_buffer():
    """Addresses from p2p_plan: at every destination the (bucket, source) runs tile the buffer
    exactly, bucket-major and source-minor, and seg_off bounds the coarse buckets."""
    from dpu_olap_b200.sharded import p2p_plan
    G, C = 4, 8
    B = G * C
    rng = np.random.default_rng(3)
    counts = torch.from_numpy(rng.integers(0, 50, size=(G, B)).astype(np.int64))
    counts[1, 5] = 0
    bases = torch.tensor([1 << 40, 2 << 40, 3 << 40, 4 << 40], dtype=torch.int64)
    plans = [p2p_plan(counts, bases, r, G) for r in range(G)]
    for dest in range(G):
        runs = []  # (start row, rows, bucket, source) of everything written into dest's buffer
        for src in range(G):
            addr = plans[src][0]
            for b in range(dest * C, (dest + 1) * C):
                start = (int(addr[b]) - int(bases[dest]))
                assert start % 8 == 0
                runs.append((start // 8, int(counts[src, b]), b, src))
        runs.sort(key=lambda r: (r[2], r[3]))
        pos = 0
        for start, rows, b, src in runs:
            assert start == pos, (dest, b, src)
            pos += rows
        seg_off, n_recv, max_recv = plans[dest][1], plans[dest][2], plans[dest][3]
        assert int(n_recv) == pos == int(seg_off[-1])
        for c in range(C):
            assert int(seg_off[c + 1] - seg_off[c]) == int(counts[:, dest * C + c].sum())
        assert int(max_recv) == max(int(counts[:, r * C:(r + 1) * C].sum()) for r in range(G))
