"""GPU parity tests of the device-resident C-ABI entry points (b2_*_dev) against the CPU oracle.

Bit-exact everywhere (integer work): filter output + per-batch chunk boundaries, sum, take and the
generator must match exactly; partition must route every row to the oracle's bucket and keep the
columns aligned; join must match as a sorted multiset of (fk, y, x) rows (join_test.cc:27-38).
"""
import hashlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.uint32).tobytes())
    return h.hexdigest()


# ---- generator ----------------------------------------------------------------------------------
def test_gen_matches_real_libstdcxx_vectors(ctx, golden):
    for name, v in golden["generator_golden"].items():
        if not isinstance(v, dict):
            continue
        span = v["hi"] - v["lo"] + 1
        if span & (span - 1):
            continue  # rejection-sampled spans are refused by the device generator (tested below)
        a = host(ctx.gen_dev([v["seed"] + 1], 1, v["n"], v["lo"], v["hi"]))
        assert list(a[:8]) == v["first"], name
        assert int(a[-1]) == v["last"], name
        assert int(a.astype(np.uint64).sum()) == v["sum"], name


def test_gen_refuses_rejection_spans(ctx):
    from dpu_olap_b200._lib import B2Error
    with pytest.raises(B2Error) as e:
        ctx.gen_dev([5], 1, 16, 10, 1009)
    assert e.value.status == 4  # B2_ERR_UNSUPPORTED


def test_gen_batches_vs_oracle(ctx):
    g = oracle.RandomArrayGenerator(42)
    seeds = [g.data_seed() for _ in range(5)]
    for batch_len in (1, 1000, 65536, 200_000):
        got = host(ctx.gen_dev(seeds, 5, batch_len)).reshape(5, batch_len)
        for b, s in enumerate(seeds):
            assert np.array_equal(got[b], oracle.gen_u32(s, batch_len)), (batch_len, b)
    lo = np.array([i << 21 for i in range(5)], dtype=np.uint32)
    hi = lo + ((1 << 21) - 1)
    got = host(ctx.gen_dev(seeds, 5, 4096, lo, hi)).reshape(5, 4096)
    for b, s in enumerate(seeds):
        assert np.array_equal(got[b], oracle.gen_u32(s, 4096, int(lo[b]), int(hi[b])))


def test_iota(ctx):
    assert np.array_equal(host(ctx.iota_dev(2**32 - 5, 10)), oracle.iota_u32(2**32 - 5, 10))


# ---- sum ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 5, 4095, 4096, 4097, 65536, 1_000_003, 8 * 1024 * 1024 + 3])
def test_sum_sizes(ctx, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    out = ctx.sum_dev(dev(a) if n else torch.empty(0, dtype=torch.int32, device="cuda"))
    assert int(out.cpu().numpy().view(np.uint64)[0]) == oracle.sum_u32(a)


def test_sum_unaligned_and_max(ctx):
    a = np.full(1_000_000, 0xFFFFFFFF, dtype=np.uint32)
    d = dev(a)
    for off in (0, 1, 2, 3, 5):
        out = ctx.sum_dev(d[off:])
        assert int(out.cpu().numpy().view(np.uint64)[0]) == oracle.sum_u32(a[off:])


def test_sum_reference_kats(ctx, golden):
    assert int(ctx.sum_dev(dev([0, 2, 3, 8, 9])).cpu().numpy().view(np.uint64)[0]) == 22  # aggr_test.cc:24-35
    g = oracle.RandomArrayGenerator(42)
    seeds = [g.data_seed() for _ in range(128)]
    col = ctx.gen_dev(seeds, 128, 65536)
    assert int(ctx.sum_dev(col).cpu().numpy().view(np.uint64)[0]) == golden["arrow_golden"]["sum_128x65536"]["sum"]


def test_sum_idempotent_relaunch(ctx):
    a = dev(np.arange(100_000, dtype=np.uint32))
    r = [int(ctx.sum_dev(a).cpu().numpy()[0]) for _ in range(3)]
    assert r[0] == r[1] == r[2] == 99_999 * 100_000 // 2


# ---- filter ---------------------------------------------------------------------------------------
def run_filter(ctx, batches, thr=1 << 30):
    nb = len(batches)
    bl = batches[0].size if nb else 0
    col = dev(np.concatenate(batches)) if nb and bl else torch.empty(0, dtype=torch.int32, device="cuda")
    out, end, total = ctx.filter_dev(col, nb, bl, thr)
    torch.cuda.synchronize()
    end = end.cpu().numpy()[:nb]
    total = int(total.cpu().numpy()[0])
    flat = host(out)[:total]
    return flat, end, total


def check_filter(ctx, batches, thr=1 << 30):
    flat, end, total = run_filter(ctx, batches, thr)
    exp = [oracle.filter_lt(b, thr) for b in batches]
    exp_end = np.cumsum([e.size for e in exp]) if batches else np.array([], dtype=np.int64)
    assert total == (int(exp_end[-1]) if len(exp_end) else 0)
    assert np.array_equal(end, exp_end)
    if exp:
        assert np.array_equal(flat, np.concatenate(exp))


def test_filter_reference_kats(ctx):
    check_filter(ctx, [np.array([0, 2, 3, 8, 9], dtype=np.uint32)])  # filter_test.cc:24-31
    keep = {5, 8, 9, 100, 270}  # filter_test.cc:33-61
    v = np.array([i if i in keep else i + (1 << 30) for i in range(4096)], dtype=np.uint32)
    flat, end, total = run_filter(ctx, [v])
    assert list(flat) == [5, 8, 9, 100, 270] and total == 5


def test_filter_longer_test_vs_arrow_digest(ctx, golden):  # filter_test.cc:63-78 + 128-batch digest
    ref = golden["arrow_golden"]["filter_128x65536"]
    g = oracle.RandomArrayGenerator(42)
    seeds = [g.data_seed() for _ in range(128)]
    col = ctx.gen_dev(seeds, 128, 65536)
    out, end, total = ctx.filter_dev(col, 128, 65536, 1 << 30)
    torch.cuda.synchronize()
    end = end.cpu().numpy()
    assert int(total.cpu()[0]) == ref["rows"]
    assert np.diff(np.concatenate([[0], end]))[:8].tolist() == ref["per_batch_first8"]
    flat = host(out)[: ref["rows"]]
    assert sha(flat) == ref["sha256"]
    assert sha(flat[: end[0]]) == ref["batch0_sha256"]


@pytest.mark.parametrize("nb,bl", [(0, 0), (1, 0), (3, 1), (1, 8191), (1, 8192), (1, 8193), (7, 1000),
                                   (5, 65536), (3, 65537), (2, 100_003), (40, 8192)])
def test_filter_shapes(ctx, nb, bl):
    rng = np.random.default_rng(nb * 1000 + bl)
    check_filter(ctx, [rng.integers(0, 2**32, size=bl, dtype=np.uint32) for _ in range(nb)])


@pytest.mark.parametrize("thr", [0, 1, 42_949_673, 429_496_730, 1 << 31, 0xFFFFFFFF])
def test_filter_selectivities(ctx, thr):  # 0 %, ~0 %, 1 %, 10 %, 50 %, ~100 %
    rng = np.random.default_rng(thr % 1000)
    batches = [rng.integers(0, 2**32, size=65536, dtype=np.uint32) for _ in range(20)]
    batches[3][:] = 0xFFFFFFFF  # a batch that selects nothing (unless thr is max) ...
    batches[4][:] = 0           # ... and one that selects everything (unless thr is 0)
    check_filter(ctx, batches, thr)


def test_filter_many_tiles_lookback_chain(ctx):
    """> 32 tiles per look-back window and thousands of tiles: every tile's prefix must chain."""
    rng = np.random.default_rng(7)
    batches = [rng.integers(0, 2**32, size=65536, dtype=np.uint32) for _ in range(512)]
    check_filter(ctx, batches)


def test_filter_ragged(ctx):
    rng = np.random.default_rng(11)
    lens = [0, 5, 8192, 0, 8193, 1, 70_000, 0, 4096, 3]
    batches = [rng.integers(0, 2**32, size=n, dtype=np.uint32) for n in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = dev(np.concatenate(batches))
    out, end, total = ctx.filter_ragged_dev(col, off, 1 << 30)
    torch.cuda.synchronize()
    exp = [oracle.filter_lt(b) for b in batches]
    assert np.array_equal(end.cpu().numpy()[: len(lens)], np.cumsum([e.size for e in exp]))
    assert np.array_equal(host(out)[: int(total.cpu()[0])], np.concatenate(exp))


def test_filter_carry_in_appends(ctx):
    """Two calls chained through d_carry_in produce one compacted result (streaming upload path)."""
    rng = np.random.default_rng(5)
    a = [rng.integers(0, 2**32, size=8192 * 3, dtype=np.uint32) for _ in range(2)]
    b = [rng.integers(0, 2**32, size=8192 * 3, dtype=np.uint32) for _ in range(3)]
    out = torch.empty(5 * 8192 * 3, dtype=torch.int32, device="cuda")
    _, end_a, tot_a = ctx.filter_dev(dev(np.concatenate(a)), 2, 8192 * 3, 1 << 30, out=out)
    _, end_b, tot_b = ctx.filter_dev(dev(np.concatenate(b)), 3, 8192 * 3, 1 << 30, out=out, carry_in=tot_a)
    torch.cuda.synchronize()
    exp = [oracle.filter_lt(x) for x in a + b]
    ends = np.concatenate([end_a.cpu().numpy()[:2], end_b.cpu().numpy()[:3]])
    assert np.array_equal(ends, np.cumsum([e.size for e in exp]))
    assert np.array_equal(host(out)[: int(tot_b.cpu()[0])], np.concatenate(exp))


def test_filter_idempotent_and_sorted_property(ctx):
    """filter(filter(x)) == filter(x); the output of a sorted column is a prefix of it."""
    col = np.sort(np.random.default_rng(3).integers(0, 2**32, size=8 * 65536, dtype=np.uint32))
    flat, _, total = run_filter(ctx, list(col.reshape(8, 65536)))
    assert np.array_equal(flat, col[:total])
    pad = (-total) % 8
    again = np.concatenate([flat, np.full(pad, 0xFFFFFFFF, dtype=np.uint32)])
    flat2, _, total2 = run_filter(ctx, list(again.reshape(8, -1)))
    assert total2 == total and np.array_equal(flat2, flat)


# ---- take -----------------------------------------------------------------------------------------
def test_take_reference_kats(ctx, golden):
    out = ctx.take_dev(dev([0, 2, 3, 8, 9]), 5, dev([0, 1, 4]), 3, 1)  # take_test.cc:24-46
    assert list(host(out)) == [0, 2, 9]
    ref = golden["arrow_golden"]["take_128x65536_8192"]  # take_test.cc:48-72
    g = oracle.RandomArrayGenerator(42)
    vseeds = [g.data_seed() for _ in range(128)]
    iseeds = [g.data_seed() for _ in range(128)]
    vals = ctx.gen_dev(vseeds, 128, 65536)
    idx = ctx.gen_dev(iseeds, 128, 8192, 0, 65535)
    out = host(ctx.take_dev(vals, 65536, idx, 8192, 128))
    assert sha(out) == ref["sha256"]
    assert [int(v) for v in out[:8]] == ref["batch0_first8"]


@pytest.mark.parametrize("nb,vl,il", [(1, 1, 1), (3, 7, 5), (4, 1000, 4096), (2, 4 << 20, 512 << 10),
                                      (5, 65536, 8190), (3, 10, 0)])
def test_take_shapes(ctx, nb, vl, il):
    rng = np.random.default_rng(nb + vl + il)
    vals = rng.integers(0, 2**32, size=(nb, vl), dtype=np.uint32)
    idx = rng.integers(0, vl, size=(nb, il), dtype=np.uint32)
    out = ctx.take_dev(dev(vals), vl, dev(idx) if il else torch.empty(0, dtype=torch.int32, device="cuda"), il, nb)
    exp = np.concatenate([oracle.take(vals[b], idx[b]) for b in range(nb)]) if il else np.array([], np.uint32)
    assert np.array_equal(host(out)[: nb * il], exp)


def test_take_ragged(ctx):
    rng = np.random.default_rng(9)
    vlens, ilens = [5, 1000, 1, 70_000], [3, 0, 10, 12_345]
    vals = [rng.integers(0, 2**32, size=n, dtype=np.uint32) for n in vlens]
    idx = [rng.integers(0, v, size=n, dtype=np.uint32) for v, n in zip(vlens, ilens)]
    voff = np.concatenate([[0], np.cumsum(vlens)])
    ioff = np.concatenate([[0], np.cumsum(ilens)])
    out = ctx.take_ragged_dev(dev(np.concatenate(vals)), voff, dev(np.concatenate(idx)), ioff)
    assert np.array_equal(host(out), np.concatenate([oracle.take(v, i) for v, i in zip(vals, idx)]))


# ---- partition ------------------------------------------------------------------------------------
def check_partition(ctx, cols, nparts, skip_bits=0):
    outs, off = ctx.partition_dev([dev(c) for c in cols], nparts, skip_bits)
    torch.cuda.synchronize()
    off = off.cpu().numpy()
    outs = [host(o) for o in outs]
    n = cols[0].size
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
    ids = oracle.partition_ids(cols[0], nparts, skip_bits)
    assert np.array_equal(np.diff(off), np.bincount(ids, minlength=nparts))
    # every output row sits in the partition its key hashes to
    got_ids = oracle.partition_ids(outs[0], nparts, skip_bits)
    assert np.array_equal(got_ids, np.repeat(np.arange(nparts, dtype=np.uint32), np.diff(off)))
    # rows are permuted as a whole: same multiset of full rows
    a = oracle.sort_rows(*cols)
    b = oracle.sort_rows(*outs)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    return outs, off


def test_partition_reference_kat(ctx):  # partition_test.cc:21-57
    outs, off = check_partition(ctx, [np.array([0, 2, 3, 8], np.uint32), np.array([100, 101, 102, 103], np.uint32)], 2)
    assert np.diff(off).tolist() == [3, 1]
    assert int(outs[0].astype(np.uint64).sum()) == 13 and int(outs[1].astype(np.uint64).sum()) == 406


@pytest.mark.parametrize("n,nparts,ncols", [(0, 4, 2), (1, 1, 1), (1000, 2, 3), (8192, 1024, 2), (8193, 32, 2),
                                            (300_000, 1024, 2), (1_000_003, 64, 3), (2_000_000, 4096, 2),
                                            (3_000_000, 1 << 14, 2)])
def test_partition_shapes(ctx, n, nparts, ncols):
    rng = np.random.default_rng(n + nparts)
    cols = [rng.integers(0, 2**32, size=n, dtype=np.uint32) for _ in range(ncols)]
    check_partition(ctx, cols, nparts)


def test_partition_large_test_balance(ctx):  # partition_test.cc:59-92 (128 x 65536, 32 partitions)
    g = oracle.RandomArrayGenerator(42)
    cols = [np.concatenate([g.uint32(65536) for _ in range(16)]) for _ in range(3)]
    _, off = check_partition(ctx, cols, 32)
    cnt = np.diff(off)
    assert np.all(np.abs(cnt - cols[0].size / 32) / (cols[0].size / 32) <= 0.1)


def test_partition_skew_and_skip_bits(ctx):
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 50, size=500_000, dtype=np.uint32)  # 50 distinct keys: heavy skew
    check_partition(ctx, [keys, np.arange(keys.size, dtype=np.uint32)], 256)
    keys = rng.integers(0, 2**32, size=200_000, dtype=np.uint32)
    check_partition(ctx, [keys, keys ^ 0x5A5A5A5A], 128, skip_bits=3)


# ---- join -----------------------------------------------------------------------------------------
def run_join(ctx, fk, y, pk, x, cap=None, ws_bytes=None, skip_bits=0):
    ws = None
    if ws_bytes is not None:
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    e = lambda a: dev(a) if len(a) else torch.empty(0, dtype=torch.int32, device="cuda")
    o_fk, o_y, o_x, rows = ctx.join_dev(e(fk), e(y), e(pk), e(x), out_capacity=cap, ws=ws, skip_bits=skip_bits)
    torch.cuda.synchronize()
    n = int(rows.cpu().numpy().view(np.uint64)[0])
    return host(o_fk)[:n], host(o_y)[:n], host(o_x)[:n], n


def check_join(ctx, fk, y, pk, x, **kw):
    exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    g_fk, g_y, g_x, n = run_join(ctx, fk, y, pk, x, cap=max(exp[0].size, 1), **kw)
    assert n == exp[0].size
    got = oracle.sort_rows(g_fk, g_y, g_x)
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)


def test_join_simple_test(ctx):  # join_test.cc:40-80
    fk = np.array([0, 2, 3, 8, 9, 10, 12, 13, 18, 19], np.uint32)
    vl = fk + 100
    pk = np.array([3, 8, 9, 0, 12, 13, 18, 19, 10, 2], np.uint32)
    vr = pk + 50
    check_join(ctx, fk, vl, pk, vr)


def test_join_large_test_vs_arrow_digest(ctx, golden):  # join_test.cc:82-121
    ref = golden["arrow_golden"]["join_128x65536"]
    g = oracle.RandomArrayGenerator(42)
    nb, bs = 128, 65536
    xs = [g.data_seed() for _ in range(nb)]
    ys = [g.data_seed() for _ in range(nb)]
    fs = [g.data_seed() for _ in range(nb)]
    x = ctx.gen_dev(xs, nb, bs)
    pk = ctx.iota_dev(0, nb * bs)
    y = ctx.gen_dev(ys, nb, bs)
    lo = np.array([i * bs for i in range(nb)], dtype=np.uint32)
    fk = ctx.gen_dev(fs, nb, bs, lo, lo + (bs - 1))
    o_fk, o_y, o_x, rows = ctx.join_dev(fk, y, pk, x)
    torch.cuda.synchronize()
    n = int(rows.cpu()[0])
    assert n == ref["rows"] == nb * bs  # join_test.cc:115-116
    got = oracle.sort_rows(host(o_fk)[:n], host(o_y)[:n], host(o_x)[:n])
    assert sha(*got) == ref["sorted_sha256"]
    assert oracle.triple_checksum(*got) == ref["checksum"]


@pytest.mark.parametrize("nl,nr", [(0, 0), (0, 10), (10, 0), (1, 1), (100, 7), (7, 100), (5000, 5000),
                                   (100_000, 30_000), (1 << 20, 1 << 20), (3_000_000, 500_000)])
def test_join_shapes_unique_build(ctx, nl, nr):
    rng = np.random.default_rng(nl * 7 + nr)
    pk = rng.permutation(np.arange(nr, dtype=np.uint32) * 3 + 1) if nr else np.array([], np.uint32)
    x = rng.integers(0, 2**32, size=nr, dtype=np.uint32)
    # half of the probe keys miss (inner join drops them; the DPU path would assert)
    fk = rng.integers(0, max(3 * nr, 1) * 2, size=nl, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=nl, dtype=np.uint32)
    check_join(ctx, fk, y, pk, x)


def test_join_duplicates_both_sides(ctx):
    rng = np.random.default_rng(2)
    pk = rng.integers(0, 2000, size=20_000, dtype=np.uint32)   # ~10 duplicates per build key
    x = np.arange(pk.size, dtype=np.uint32)
    fk = rng.integers(0, 2500, size=30_000, dtype=np.uint32)
    y = np.arange(fk.size, dtype=np.uint32) + 7
    check_join(ctx, fk, y, pk, x)


def test_join_heavy_skew_oversized_partition(ctx):
    """One build key repeated 20000 times: its partition exceeds the shared-memory table and is
    built in chunks; extreme key values exercise the empty-slot marker."""
    pk = np.concatenate([np.full(20_000, 12345, np.uint32), np.array([0, 0xFFFFFFFF, 0xFFFFFFFE, 1], np.uint32)])
    x = np.arange(pk.size, dtype=np.uint32)
    fk = np.array([12345, 0xFFFFFFFF, 0, 5, 12345, 0xFFFFFFFE], np.uint32)
    y = np.arange(fk.size, dtype=np.uint32)
    check_join(ctx, fk, y, pk, x)


@pytest.mark.parametrize("nl,nr,domain,thr", [(0, 10, 10, None), (10, 0, 10, 5), (5000, 5000, 5000, None),
                                              (300_000, 100_000, 100_000, 1 << 30), (200_000, 50_000, 2000, 1 << 31),
                                              (1 << 21, 1 << 21, 1 << 21, 1 << 30), (30_000, 20_000, 50, None)])
def test_join_aggregate_fused_pipeline(ctx, nl, nr, domain, thr):
    """b2_join_aggr_u32_dev == filter (probe side) -> join -> count / sum y / sum x of the oracle."""
    rng = np.random.default_rng(nl + nr + domain)
    pk = rng.integers(0, max(domain, 1), size=nr, dtype=np.uint32)       # duplicates when domain < nr
    x = rng.integers(0, 2**32, size=nr, dtype=np.uint32)
    fk = rng.integers(0, max(domain, 1) + domain // 4 + 1, size=nl, dtype=np.uint32)  # some probes miss
    y = rng.integers(0, 2**32, size=nl, dtype=np.uint32)
    if domain == 50:
        pk, x, fk, y = pk[:2000], x[:2000], fk[:3000], y[:3000]          # keep the quadratic output small
    e = lambda a: dev(a) if len(a) else torch.empty(0, dtype=torch.int32, device="cuda")
    out = ctx.join_aggr_dev(e(fk), e(y), e(pk), e(x), y_threshold=thr)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    exp = oracle.join_aggr(fk, y, pk, x, thr)
    assert {"rows": int(got[0]), "sum_y": int(got[1]), "sum_x": int(got[2])} == exp
    # and the unfused path agrees: materialise the join, then aggregate its columns
    if thr is None and len(fk) and len(pk):
        o_fk, o_y, o_x, rows = ctx.join_dev(e(fk), e(y), e(pk), e(x), out_capacity=max(exp["rows"], 1))
        torch.cuda.synchronize()
        n = int(rows.cpu()[0])
        assert n == exp["rows"]
        assert int(host(o_x)[:n].astype(np.uint64).sum(dtype=np.uint64)) == exp["sum_x"]


def test_join_capacity_overflow_reports_true_count(ctx):
    pk = np.zeros(100, np.uint32)
    fk = np.zeros(100, np.uint32)
    _, _, _, n = run_join(ctx, fk, fk, pk, pk, cap=10)
    assert n == 10_000  # true match count although only 10 rows were written


def test_join_sliced_small_workspace(ctx):
    """With a workspace below b2_join_ws_bytes the join runs in hash-space slices."""
    rng = np.random.default_rng(4)
    nr = nl = 1 << 21
    pk = rng.permutation(nr).astype(np.uint32)
    x = rng.integers(0, 2**32, size=nr, dtype=np.uint32)
    fk = rng.integers(0, nr, size=nl, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=nl, dtype=np.uint32)
    full, small = ctx.join_ws_bytes(nl, nr), ctx.join_min_ws_bytes(nl, nr)
    assert small < full
    check_join(ctx, fk, y, pk, x, ws_bytes=(full + small) // 3)
    from dpu_olap_b200._lib import B2Error
    with pytest.raises(B2Error) as e:
        run_join(ctx, fk, y, pk, x, ws_bytes=small // 2)
    assert e.value.status == 5  # B2_ERR_WORKSPACE


def test_join_skip_bits_matches_shuffle_routing(ctx):
    """Sharded join building blocks: route rows by b2_shuffle_partition (top hash bits), join each
    rank's share with hash_skip_bits, and the union equals the unsharded join."""
    from dpu_olap_b200.ops import join_dest_rank
    rng = np.random.default_rng(8)
    nr = nl = 300_000
    pk = rng.permutation(nr).astype(np.uint32)
    x = rng.integers(0, 2**32, size=nr, dtype=np.uint32)
    fk = rng.integers(0, nr, size=nl, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=nl, dtype=np.uint32)
    G = 4
    lp, loff = ctx.shuffle_partition_dev(dev(fk), dev(y), G)
    rp, roff = ctx.shuffle_partition_dev(dev(pk), dev(x), G)
    torch.cuda.synchronize()
    loff, roff = loff.cpu().numpy(), roff.cpu().numpy()
    assert loff[-1] == nl and roff[-1] == nr
    lkeys = (lp.cpu().numpy().view(np.uint64) & 0xFFFFFFFF).astype(np.uint32)
    for r in range(G):
        seg = lkeys[loff[r]:loff[r + 1]]
        assert all(join_dest_rank(int(k), G) == r for k in seg[:50])
        assert np.array_equal(oracle.partition_ids(seg, G), np.full(seg.size, r, np.uint32))
    parts = []
    for r in range(G):
        l_r, r_r = lp[loff[r]:loff[r + 1]], rp[roff[r]:roff[r + 1]]
        o = ctx.join_pairs_dev(l_r, r_r, out_capacity=max(l_r.numel(), 1), skip_bits=2)
        torch.cuda.synchronize()
        n = int(o[3].cpu()[0])
        parts.append([host(t)[:n] for t in o[:3]])
    got = oracle.sort_rows(*[np.concatenate([p[c] for p in parts]) for c in range(3)])
    exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)


# ---- fused multi-GPU shuffle, emulated with virtual ranks on one GPU -------------------------------
@pytest.fixture(params=[0, 1, 2], ids=["lines", "bulk", "bulk-budget"])
def peer_kernel(ctx, request):
    """Both peer scatter kernels (whole 128-byte lines stored by the threads / whole sectors through the
    copy engine), the second also with a CTA budget (3 persistent CTAs walk all work units)."""
    from dpu_olap_b200._lib import TUNE_PEER_SCATTER_CTAS, TUNE_PEER_SCATTER_KERNEL
    ctx.set_tunable(TUNE_PEER_SCATTER_KERNEL, 1 if request.param else 0)
    ctx.set_tunable(TUNE_PEER_SCATTER_CTAS, 3 if request.param == 2 else 0)
    yield ctx
    ctx.set_tunable(TUNE_PEER_SCATTER_KERNEL, 0)
    ctx.set_tunable(TUNE_PEER_SCATTER_CTAS, 0)


@pytest.mark.parametrize("G,rows_per_rank", [(2, 70_000), (4, 300_000), (8, 40_000), (1, 50_000)])
def test_p2p_shuffle_and_segmented_join_virtual_ranks(peer_kernel, G, rows_per_rank):
    """Every virtual rank counts, scatters straight into the (local stand-ins for the) peers'
    receive buffers, and joins what it received with the segmented join; the union of the G
    results must equal the oracle join of the whole input, and every received row must belong to
    its rank and sit in its coarse bucket."""
    ctx = peer_kernel
    from dpu_olap_b200.sharded import log2_exact, p2p_plan
    BITS = 10
    skip = log2_exact(G)
    seg_bits = BITS - skip
    C = 1 << seg_bits
    rng = np.random.default_rng(G)
    n = G * rows_per_rank
    pk = rng.permutation(np.arange(n, dtype=np.uint32) * 7 + 3)            # unique build keys
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    fk = pk[rng.integers(0, n, size=n)]                                      # every probe row matches
    fk[:100] = 1                                                             # ... except these
    y = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    sl = lambda a, r: a[r * rows_per_rank:(r + 1) * rows_per_rank]
    cap = rows_per_rank * 2 + 1024
    recv = [[torch.zeros(cap, dtype=torch.int64, device="cuda") for _ in range(G)] for _ in range(2)]
    peers = [torch.tensor([t.data_ptr() for t in recv[s]], dtype=torch.int64, device="cuda") for s in range(2)]
    d = {r: [dev(sl(a, r)) for a in (fk, y, pk, x)] for r in range(G)}
    wss = {(r, s): torch.empty(ctx.shuffle_p2p_ws_bytes(rows_per_rank, BITS) + 256, dtype=torch.uint8, device="cuda")
           for r in range(G) for s in range(2)}
    offs = {}
    for r in range(G):
        offs[r, 0] = ctx.shuffle_p2p_count_dev(d[r][0], BITS, wss[r, 0])
        offs[r, 1] = ctx.shuffle_p2p_count_dev(d[r][2], BITS, wss[r, 1])
    allc = [torch.stack([offs[r, s][1:] - offs[r, s][:-1] for r in range(G)]) for s in range(2)]
    # counts agree with the oracle's routing
    assert np.array_equal(allc[0][0].cpu().numpy(), np.bincount(oracle.partition_ids(sl(fk, 0), 1 << BITS), minlength=1 << BITS))
    plans = {(r, s): p2p_plan(allc[s], peers[s], r, G) for r in range(G) for s in range(2)}
    for r in range(G):
        ctx.shuffle_p2p_scatter_dev(d[r][0], d[r][1], BITS, plans[r, 0][0], wss[r, 0])
        ctx.shuffle_p2p_scatter_dev(d[r][2], d[r][3], BITS, plans[r, 1][0], wss[r, 1])
    torch.cuda.synchronize()
    got = [[], [], []]
    for r in range(G):
        nl, nr = int(plans[r, 0][2]), int(plans[r, 1][2])
        lrecv, rrecv = recv[0][r][:nl], recv[1][r][:nr]
        # received rows: right rank, right coarse bucket
        k = (lrecv.cpu().numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        ids = oracle.partition_ids(k, 1 << BITS)
        assert np.all(ids >> seg_bits == r) if G > 1 else True
        seg = plans[r, 0][1].cpu().numpy()
        assert np.array_equal(np.bincount(ids & (C - 1), minlength=C), np.diff(seg))
        assert np.all(np.diff(ids) >= 0)  # bucket-major
        o_fk, o_y, o_x, rows = ctx.join_pairs_seg_dev(lrecv, plans[r, 0][1], rrecv, plans[r, 1][1], seg_bits,
                                                      out_capacity=max(nl, 1), skip_bits=skip)
        torch.cuda.synchronize()
        m = int(rows.cpu().numpy().view(np.uint64)[0])
        for i, t in enumerate((o_fk, o_y, o_x)):
            got[i].append(host(t)[:m])
    got = oracle.sort_rows(*[np.concatenate(g) for g in got])
    exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    assert got[0].size == exp[0].size
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)


def test_segmented_join_large_needs_fine_pass(ctx):
    """2^22 rows per side in 128 coarse buckets: the fine pass (2^3 per bucket) must run."""
    BITS, G = 10, 8
    n = 1 << 25
    rng = np.random.default_rng(11)
    pk = rng.permutation(n).astype(np.uint32)
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    fk = rng.integers(0, n, size=n, dtype=np.uint32)
    y = np.arange(n, dtype=np.uint32)
    # keep only the rows virtual rank 5 of 8 would own, pre-grouped by the scatter kernel itself
    keep_l = oracle.partition_ids(fk, G) == 5
    keep_r = oracle.partition_ids(pk, G) == 5
    fk, y, pk, x = fk[keep_l], y[keep_l], pk[keep_r], x[keep_r]
    out = {}
    for side, (k, v) in enumerate(((fk, y), (pk, x))):
        ws = torch.empty(ctx.shuffle_p2p_ws_bytes(k.size, BITS) + 256, dtype=torch.uint8, device="cuda")
        dk, dv = dev(k), dev(v)
        off = ctx.shuffle_p2p_count_dev(dk, BITS, ws)
        buf = torch.zeros(k.size, dtype=torch.int64, device="cuda")
        addr = buf.data_ptr() + 8 * off[:-1]
        ctx.shuffle_p2p_scatter_dev(dk, dv, BITS, addr.contiguous(), ws)
        out[side] = (buf, off[5 * 128: 6 * 128 + 1] - off[5 * 128])
    o_fk, o_y, o_x, rows = ctx.join_pairs_seg_dev(out[0][0], out[0][1].contiguous(), out[1][0],
                                                  out[1][1].contiguous(), 7, out_capacity=fk.size, skip_bits=3)
    torch.cuda.synchronize()
    m = int(rows.cpu().numpy().view(np.uint64)[0])
    got = oracle.sort_rows(host(o_fk)[:m], host(o_y)[:m], host(o_x)[:m])
    exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    assert m == exp[0].size
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)


def test_filter_unaligned_column_takes_the_non_tma_path(ctx):
    """A column that starts 4, 8 or 12 bytes off a 16-byte boundary cannot be fetched with TMA bulk
    copies; the kernel then loads those tiles with ordinary loads. Same result either way."""
    rng = np.random.default_rng(17)
    base = rng.integers(0, 2**32, size=3 * 65536 + 8, dtype=np.uint32)
    d = dev(base)
    for shift in (1, 2, 3):
        col = d[shift: shift + 3 * 65536]
        assert col.data_ptr() % 16 == 4 * shift
        out, end, total = ctx.filter_dev(col, 3, 65536, 1 << 30)
        torch.cuda.synchronize()
        exp = [oracle.filter_lt(base[shift + b * 65536: shift + (b + 1) * 65536]) for b in range(3)]
        n = int(total.cpu()[0])
        assert n == sum(e.size for e in exp)
        assert np.array_equal(host(out)[:n], np.concatenate(exp))
        assert end.cpu().numpy().tolist() == np.cumsum([e.size for e in exp]).tolist()


def test_pinned_host_memory_entry_points(ctx):
    import ctypes as C
    lib = ctx._lib
    p = C.c_void_p()
    assert lib.b2_host_alloc_pinned(1 << 20, C.byref(p)) == 0 and p.value
    a = np.ctypeslib.as_array((C.c_uint32 * (1 << 18)).from_address(p.value))
    a[:] = np.arange(1 << 18, dtype=np.uint32)
    t = torch.from_numpy(a)
    assert int(t.cuda().to(torch.int64).sum()) == int(a.astype(np.int64).sum())
    del t, a
    assert lib.b2_host_free_pinned(p) == 0
    b = np.arange(1 << 18, dtype=np.uint32)
    assert lib.b2_host_register(b.ctypes.data, b.nbytes) == 0
    assert int(torch.from_numpy(b).cuda().to(torch.int64).sum()) == int(b.astype(np.int64).sum())
    assert lib.b2_host_unregister(b.ctypes.data) == 0
    assert lib.b2_host_unregister(b.ctypes.data) == 0  # idempotent


# ---- out-of-bounds canaries (compute-sanitizer is not available on the GPU pool) ---------------------
GUARD = 4096  # elements of sentinel on both sides of every output


def _guarded(n, dtype, fill):
    t = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device="cuda")
    return t, t[GUARD: GUARD + n]


def _guards_intact(t, n, fill):
    return bool((t[:GUARD] == fill).all()) and bool((t[GUARD + n:] == fill).all())


def test_kernels_never_write_outside_their_outputs(ctx):
    """Every output buffer sits between sentinel guards; after each operator the guards must be
    untouched (exact-size outputs, ragged / odd sizes, 100 % selectivity, partial tiles)."""
    rng = np.random.default_rng(31)
    S32, S64 = -559038737, -81985529216486896
    # filter: 100 % selected, odd batch length, output capacity exactly n
    nb, bl = 5, 8191 + 4096
    col = dev(rng.integers(0, 2**32, size=nb * bl, dtype=np.uint32))
    g_out, out = _guarded(nb * bl, torch.int32, S32)
    g_end, end = _guarded(nb, torch.int64, S64)
    g_tot, tot = _guarded(1, torch.int64, S64)
    ws_n = ctx.filter_ws_bytes(nb, bl)
    g_ws, ws = _guarded(ws_n, torch.uint8, 0x5A)
    ctx.filter_dev(col, nb, bl, 0xFFFFFFFF, out=out, batch_end=end, total=tot, ws=ws)
    torch.cuda.synchronize()
    assert int(tot.cpu()[0]) == int((host(col) < 0xFFFFFFFF).sum())
    assert _guards_intact(g_out, nb * bl, S32) and _guards_intact(g_end, nb, S64)
    assert _guards_intact(g_tot, 1, S64) and _guards_intact(g_ws, ws_n, 0x5A)
    # take: odd sizes
    vals = dev(rng.integers(0, 2**32, size=3 * 1001, dtype=np.uint32))
    idx = dev(rng.integers(0, 1001, size=3 * 77, dtype=np.uint32))
    g_o, o = _guarded(3 * 77, torch.int32, S32)
    ctx.take_dev(vals, 1001, idx, 77, 3, out=o)
    torch.cuda.synchronize()
    assert _guards_intact(g_o, 3 * 77, S32)
    # join: exact output capacity, duplicates on the build side
    n = 50_001
    pk = rng.integers(0, 20_000, size=n, dtype=np.uint32)
    fk = rng.integers(0, 25_000, size=n, dtype=np.uint32)
    x, y = rng.integers(0, 2**32, size=n, dtype=np.uint32), rng.integers(0, 2**32, size=n, dtype=np.uint32)
    exp_rows = oracle.join(fk, y, pk, x)[0].size
    guards = [_guarded(exp_rows, torch.int32, S32) for _ in range(3)]
    g_r, r = _guarded(1, torch.int64, S64)
    ctx.join_dev(dev(fk), dev(y), dev(pk), dev(x), out_capacity=exp_rows, outs=[g[1] for g in guards], out_rows=r)
    torch.cuda.synchronize()
    assert int(r.cpu()[0]) == exp_rows
    assert all(_guards_intact(g[0], exp_rows, S32) for g in guards) and _guards_intact(g_r, 1, S64)
    # same join with HALF the needed capacity: the true count is reported, nothing beyond capacity is written
    cap = exp_rows // 2
    guards = [_guarded(cap, torch.int32, S32) for _ in range(3)]
    ctx.join_dev(dev(fk), dev(y), dev(pk), dev(x), out_capacity=cap, outs=[g[1] for g in guards], out_rows=r)
    torch.cuda.synchronize()
    assert int(r.cpu()[0]) == exp_rows and all(_guards_intact(g[0], cap, S32) for g in guards)
    # fused shuffle scatter into an exactly sized receive buffer (line carry: partial first/last lines)
    m = 123_457
    k, v = dev(rng.integers(0, 2**32, size=m, dtype=np.uint32)), dev(np.arange(m, dtype=np.uint32))
    sws = torch.empty(ctx.shuffle_p2p_ws_bytes(m, 10) + 256, dtype=torch.uint8, device="cuda")
    off = ctx.shuffle_p2p_count_dev(k, 10, sws)
    g_b, buf = _guarded(m, torch.int64, S64)
    for shift in (0, 1):  # receive buffer 16-byte aligned and 8 bytes off
        b2 = buf if shift == 0 else g_b[GUARD + 1: GUARD + 1 + m - 1]
        cnt = m if shift == 0 else m - 1
        if shift == 1:
            k2, v2 = k[: m - 1].contiguous(), v[: m - 1].contiguous()
            off = ctx.shuffle_p2p_count_dev(k2, 10, sws)
        else:
            k2, v2 = k, v
        g_b.fill_(S64)
        addr = (b2.data_ptr() + 8 * off[:-1]).contiguous()
        ctx.shuffle_p2p_scatter_dev(k2, v2, 10, addr, sws)
        torch.cuda.synchronize()
        lo = GUARD + shift
        assert bool((g_b[:lo] == S64).all()) and bool((g_b[lo + cnt:] == S64).all())
        got = np.sort(b2.cpu().numpy().view(np.uint64))
        expp = np.sort(host(k2).astype(np.uint64) | (host(v2).astype(np.uint64) << np.uint64(32)))
        assert np.array_equal(got, expp)


def test_nullable_and_fused_kernels_never_write_outside_their_outputs(ctx):
    """The same guard check for the nullable filter / take / aggregate kernels and the fused
    join -> aggregate pipeline (whose only outputs are 24 bytes and its workspace)."""
    rng = np.random.default_rng(32)
    S32, S64 = -559038737, -81985529216486896
    nb, bl = 3, 4096 + 36
    v = rng.integers(0, 2**32, size=nb * bl, dtype=np.uint32)
    valid = rng.random(nb * bl) > 0.05
    bits = torch.from_numpy(oracle.pack_bits(valid)).cuda()
    g_out, out = _guarded(int(valid.sum()), torch.int32, S32)   # exactly the selected rows at 100 % selectivity
    g_end, end = _guarded(nb, torch.int64, S64)
    g_tot, tot = _guarded(1, torch.int64, S64)
    ws_n = ctx.filter_ws_bytes(nb, bl)
    g_ws, ws = _guarded(ws_n, torch.uint8, 0x5A)
    ctx.filter_nullable_dev(dev(v), bits, nb, bl, 0xFFFFFFFF, out=out, batch_end=end, total=tot, ws=ws)
    torch.cuda.synchronize()
    assert int(tot.cpu()[0]) == int((valid & (v < 0xFFFFFFFF)).sum())
    # (a value equal to 0xFFFFFFFF is not selected, so the output may be a few rows short of `valid.sum()`)
    assert _guards_intact(g_out, int(valid.sum()), S32) and _guards_intact(g_end, nb, S64)
    assert _guards_intact(g_tot, 1, S64) and _guards_intact(g_ws, ws_n, 0x5A)
    # take: 3 x 77 outputs -> a result bitmap of ceil(231 / 32) words
    vals = rng.integers(0, 2**32, size=3 * 1001, dtype=np.uint32)
    idx = rng.integers(0, 1001, size=3 * 77, dtype=np.uint32)
    vv, iv = rng.random(vals.size) > 0.3, rng.random(idx.size) > 0.3
    g_o, o = _guarded(3 * 77, torch.int32, S32)
    nwords = (3 * 77 + 31) // 32
    g_b, ob = _guarded(nwords * 4, torch.uint8, 0x5A)
    ctx.take_nullable_dev(dev(vals), torch.from_numpy(oracle.pack_bits(vv)).cuda(), 1001, dev(idx),
                          torch.from_numpy(oracle.pack_bits(iv)).cuda(), 77, 3, out=o, out_valid=ob)
    torch.cuda.synchronize()
    assert _guards_intact(g_o, 3 * 77, S32) and _guards_intact(g_b, nwords * 4, 0x5A)
    # aggregates: a 24-byte result
    g_a, a = _guarded(3, torch.int64, S64)
    ctx.aggr_dev(dev(v), bits, out=a)
    torch.cuda.synchronize()
    assert _guards_intact(g_a, 3, S64)
    # fused join -> aggregate: result and workspace
    n = 60_001
    pk = rng.integers(0, 30_000, size=n, dtype=np.uint32)
    fk = rng.integers(0, 35_000, size=n, dtype=np.uint32)
    x, y = rng.integers(0, 2**32, size=n, dtype=np.uint32), rng.integers(0, 2**32, size=n, dtype=np.uint32)
    wsn = ctx.join_ws_bytes(n, n) + 256
    g_w, w = _guarded(wsn, torch.uint8, 0x5A)
    g_r, r = _guarded(3, torch.int64, S64)
    ctx.join_aggr_dev(dev(fk), dev(y), dev(pk), dev(x), y_threshold=1 << 31, ws=w, out=r)
    torch.cuda.synchronize()
    got = r.cpu().numpy().view(np.uint64)
    assert {"rows": int(got[0]), "sum_y": int(got[1]), "sum_x": int(got[2])} == oracle.join_aggr(fk, y, pk, x, 1 << 31)
    assert _guards_intact(g_r, 3, S64) and _guards_intact(g_w, wsn, 0x5A)


@pytest.mark.parametrize("n,thr", [(0, 5), (1, 0), (7, 2**31), (4097, 1 << 30), (1_000_003, 42_949_673),
                                   (8 * 1024 * 1024 + 3, 0xFFFFFFFF)])
def test_fused_filter_sum(ctx, n, thr):
    """filter(v < thr) -> sum in one pass equals summing the filter's output (and counting it)."""
    rng = np.random.default_rng(n % 1000 + 7)
    a = rng.integers(0, 2**32, size=n + 1, dtype=np.uint32)[1:]   # 4-byte-aligned but not 16
    d = dev(np.concatenate([[0], a]))[1:]
    s, c = ctx.sum_lt_dev(d, thr)
    torch.cuda.synchronize()
    sel = oracle.filter_lt(a, thr) if n else np.empty(0, np.uint32)
    assert int(s.cpu().numpy().view(np.uint64)[0]) == oracle.sum_u32(sel) if sel.size else int(s.cpu()[0]) == 0
    assert int(c.cpu()[0]) == sel.size
    # the plain sum is unaffected by the template
    p = ctx.sum_dev(d)
    assert int(p.cpu().numpy().view(np.uint64)[0]) == (oracle.sum_u32(a) if n else 0)


def test_second_device_in_the_same_process(ctx):
    """Function attributes (opt-in shared memory) are per device: a context on another GPU of the
    same process must work after the first one has launched every kernel."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dpu_olap_b200.ops import Context
    rng = np.random.default_rng(12)
    v = rng.integers(0, 2**32, size=4 * 65536, dtype=np.uint32)
    pk = rng.permutation(200_000).astype(np.uint32)
    fk = rng.integers(0, 200_000, size=300_000, dtype=np.uint32)
    ctx.filter_dev(dev(v), 4, 65536, 1 << 30)          # device 0 first
    ctx.join_dev(dev(fk), dev(fk), dev(pk), dev(pk))
    torch.cuda.synchronize()
    with torch.cuda.device(1):
        c1 = Context(1)
        d1 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to("cuda:1")
        out, _, total = c1.filter_dev(d1(v), 4, 65536, 1 << 30)
        o_fk, o_y, o_x, rows = c1.join_dev(d1(fk), d1(fk), d1(pk), d1(pk))
        torch.cuda.synchronize()
        n = int(total.cpu()[0])
        assert np.array_equal(host(out)[:n], oracle.filter_lt(v))
        assert int(rows.cpu()[0]) == fk.size
        c1.close()


@pytest.mark.parametrize("G,rows_per_rank", [(2, 60_000), (8, 30_000), (1, 20_000)])
def test_p2p_plan_kernel_and_capacity_join_virtual_ranks(ctx, G, rows_per_rank):
    """The device-side plan (b2_shuffle_p2p_plan_dev) must equal sharded.p2p_plan, and the whole step
    — count, plan, scatter, join of the receive BUFFERS by capacity — must equal the oracle join with
    nothing read back in between. A capacity that is too small must call the exchange off on the
    device: nothing stored, ~0 rows."""
    from dpu_olap_b200.sharded import log2_exact, p2p_plan
    BITS = 10
    B = 1 << BITS
    skip = log2_exact(G)
    seg_bits = BITS - skip
    rng = np.random.default_rng(100 + G)
    n = G * rows_per_rank
    pk = rng.permutation(np.arange(n, dtype=np.uint32) * 5 + 1)
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    fk = pk[rng.integers(0, n, size=n)]
    fk[:50] = 2  # no match
    y = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    sl = lambda a, r: a[r * rows_per_rank:(r + 1) * rows_per_rank]
    d = {r: [dev(sl(a, r)) for a in (fk, y, pk, x)] for r in range(G)}
    wss = {(r, s): torch.empty(ctx.shuffle_p2p_ws_bytes(rows_per_rank, BITS) + 256, dtype=torch.uint8, device="cuda")
           for r in range(G) for s in range(2)}
    all_off = torch.zeros((G, 2, B + 1), dtype=torch.int64, device="cuda")
    for r in range(G):
        ctx.shuffle_p2p_count_dev(d[r][0], BITS, wss[r, 0], all_off[r, 0])
        ctx.shuffle_p2p_count_dev(d[r][2], BITS, wss[r, 1], all_off[r, 1])
    off_ptrs = [torch.tensor([all_off[s, side].data_ptr() for s in range(G)], dtype=torch.int64, device="cuda")
                for side in range(2)]
    for cap in (rows_per_rank * 2 + 1024, rows_per_rank // 2):
        recv = [[torch.full((max(cap, 1),), -1, dtype=torch.int64, device="cuda") for _ in range(G)] for _ in range(2)]
        peers = [torch.tensor([t.data_ptr() for t in recv[s]], dtype=torch.int64, device="cuda") for s in range(2)]
        addr = {(r, s): torch.empty(B, dtype=torch.int64, device="cuda") for r in range(G) for s in range(2)}
        seg = {(r, s): torch.empty((1 << seg_bits) + 1, dtype=torch.int64, device="cuda") for r in range(G) for s in range(2)}
        info = {r: torch.zeros((2, 3), dtype=torch.int64, device="cuda") for r in range(G)}
        for r in range(G):
            ctx.shuffle_p2p_plan_dev(off_ptrs[0], peers[0], r, G, BITS, cap, addr[r, 0], seg[r, 0], info[r][0])
            ctx.shuffle_p2p_plan_dev(off_ptrs[1], peers[1], r, G, BITS, cap, addr[r, 1], seg[r, 1], info[r][1],
                                     prev_abort=info[r][0, 2:3])
        for r in range(G):
            ctx.shuffle_p2p_scatter_dev(d[r][0], d[r][1], BITS, addr[r, 0], wss[r, 0], abort=info[r][0, 2:3])
            ctx.shuffle_p2p_scatter_dev(d[r][2], d[r][3], BITS, addr[r, 1], wss[r, 1], abort=info[r][1, 2:3])
        torch.cuda.synchronize()
        counts = [all_off[:, s, 1:] - all_off[:, s, :-1] for s in range(2)]
        overflow = cap < rows_per_rank
        for r in range(G):
            for s in range(2):
                t_addr, t_seg, t_n, t_max = p2p_plan(counts[s].contiguous(), peers[s], r, G)
                assert torch.equal(addr[r, s], t_addr)
                h = info[r][s].cpu().tolist()
                assert h[1] == int(t_max)
                if not overflow:
                    assert torch.equal(seg[r, s], t_seg) and h[0] == int(t_n) and h[2] == 0
                else:
                    assert h[2] == 1 and h[0] == 0 and int(seg[r, s].abs().sum()) == 0
        got = [[], [], []]
        for r in range(G):
            nr_expected = n // G
            ws = torch.empty(ctx.join_seg_cap_ws_bytes(max(cap, 1), max(cap, 1), nr_expected, skip, seg_bits) + 256,
                             dtype=torch.uint8, device="cuda")
            outs = [torch.empty(max(cap, 1), dtype=torch.int32, device="cuda") for _ in range(3)]
            rows = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.join_pairs_seg_cap_dev(recv[0][r], seg[r, 0], recv[1][r], seg[r, 1], nr_expected, seg_bits,
                                       out_capacity=max(cap, 1), skip_bits=skip, ws=ws, outs=outs, out_rows=rows,
                                       abort=info[r][1, 2:3])
            torch.cuda.synchronize()
            m = int(rows.cpu().numpy().view(np.uint64)[0])
            if overflow:
                assert m == 0xFFFFFFFFFFFFFFFF
                assert all(int((t != -1).sum()) == 0 for t in (recv[0][r], recv[1][r]))  # nothing was stored
                agg = torch.zeros(3, dtype=torch.int64, device="cuda")
                ctx.join_aggr_pairs_seg_cap_dev(recv[0][r], seg[r, 0], recv[1][r], seg[r, 1], nr_expected, seg_bits,
                                                skip_bits=skip, ws=ws, out=agg, abort=info[r][1, 2:3])
                assert int(agg.cpu().numpy().view(np.uint64)[0]) == 0xFFFFFFFFFFFFFFFF
                continue
            for i, t in enumerate(outs):
                got[i].append(host(t)[:m])
            # the fused join -> aggregate pipeline over the same receive buffers (plain and with the
            # predicate on the left payload evaluated by the probe kernel), phased like the sharded step
            for thr in (None, 1 << 30):
                agg = torch.full((3,), -1, dtype=torch.int64, device="cuda")
                for bits in (1, 2, 4):
                    ctx.join_aggr_pairs_seg_cap_dev(recv[0][r], seg[r, 0], recv[1][r], seg[r, 1], nr_expected, seg_bits,
                                                    skip_bits=skip, ws=ws, out=agg, y_threshold=thr,
                                                    abort=info[r][1, 2:3], phases=bits)
                torch.cuda.synchronize()
                a = agg.cpu().numpy().view(np.uint64)
                keep = np.ones(m, bool) if thr is None else host(outs[1])[:m] < thr
                assert int(a[0]) == int(keep.sum())
                assert int(a[1]) == int(host(outs[1])[:m][keep].astype(np.uint64).sum(dtype=np.uint64))
                assert int(a[2]) == int(host(outs[2])[:m][keep].astype(np.uint64).sum(dtype=np.uint64))
        if not overflow:
            # the same exchange with the predicate y < 2^30 pushed in front of the "link" (count_lt / scatter_lt):
            # only the surviving probe rows are counted, planned for and stored; the aggregate of the join over
            # what arrived equals the filtered aggregate of the full join
            thr = 1 << 30
            off2 = torch.zeros((G, B + 1), dtype=torch.int64, device="cuda")
            for r in range(G):
                ctx.shuffle_p2p_count_dev(d[r][0], BITS, wss[r, 0], off2[r], val=d[r][1], val_lt=thr)
            sent = (off2[:, 1:] - off2[:, :-1]).sum().item()
            assert sent == int((y < thr).sum())
            optr = torch.tensor([off2[s].data_ptr() for s in range(G)], dtype=torch.int64, device="cuda")
            recv2 = [torch.full((cap,), -1, dtype=torch.int64, device="cuda") for _ in range(G)]
            peers2 = torch.tensor([t.data_ptr() for t in recv2], dtype=torch.int64, device="cuda")
            tot = np.zeros(3, np.uint64)
            addr2, seg2, info2 = {}, {}, {}
            for r in range(G):
                addr2[r] = torch.empty(B, dtype=torch.int64, device="cuda")
                seg2[r] = torch.empty((1 << seg_bits) + 1, dtype=torch.int64, device="cuda")
                info2[r] = torch.zeros(3, dtype=torch.int64, device="cuda")
                ctx.shuffle_p2p_plan_dev(optr, peers2, r, G, BITS, cap, addr2[r], seg2[r], info2[r])
            for r in range(G):
                ctx.shuffle_p2p_scatter_dev(d[r][0], d[r][1], BITS, addr2[r], wss[r, 0], abort=info2[r][2:3], val_lt=thr)
            for r in range(G):
                agg = torch.zeros(3, dtype=torch.int64, device="cuda")
                ws = torch.empty(ctx.join_seg_cap_ws_bytes(cap, cap, n // G, skip, seg_bits) + 256, dtype=torch.uint8,
                                 device="cuda")
                ctx.join_aggr_pairs_seg_cap_dev(recv2[r], seg2[r], recv[1][r], seg[r, 1], n // G, seg_bits,
                                                skip_bits=skip, ws=ws, out=agg, abort=info2[r][2:3])
                torch.cuda.synchronize()
                tot += agg.cpu().numpy().view(np.uint64)
                arrived = recv2[r].cpu().numpy()
                arrived = arrived[arrived != -1].view(np.uint32).reshape(-1, 2)
                assert arrived.shape[0] == int(info2[r][0].item()) and bool((arrived[:, 1] < thr).all())
            e_fk, e_y, e_x = oracle.join(fk, y, pk, x)
            keep = e_y < thr
            assert int(tot[0]) == int(keep.sum())
            assert int(tot[1]) == int(e_y[keep].astype(np.uint64).sum(dtype=np.uint64))
            assert int(tot[2]) == int(e_x[keep].astype(np.uint64).sum(dtype=np.uint64))
            got = oracle.sort_rows(*[np.concatenate(g) for g in got])
            exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
            assert got[0].size == exp[0].size
            for a, b in zip(got, exp):
                assert np.array_equal(a, b)
