"""GPU parity tests of the reference-facing operator classes (FilterGpu, SumGpu, TakeGpu, JoinGpu,
PartitionGpu — the siblings of the reference's *Dpu classes) and of the host-buffer C-ABI entry
points they call. The cases are the reference's own GoogleTests (host/*/*_test.cc), which compare
the device operator with Arrow on the same input; here the CPU oracle stands in for Arrow (it is
pinned to Arrow's outputs in tests/test_oracle.py).
"""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ops():
    from dpu_olap_b200 import ops as o
    return o


def gen_batches(nb, bl, lo=0, hi=0xFFFFFFFF, g=None):
    g = g or oracle.RandomArrayGenerator(42)
    return oracle.make_random_batches(g, nb, bl, lo, hi)


# ---- FilterTest (host/filter/filter_test.cc) -------------------------------------------------------
def test_filter_simple_test(ctx, ops):  # :24-31
    f = ops.FilterGpu(ctx, [np.array([0, 2, 3, 8, 9], np.uint32)])
    f.Prepare()
    assert f.Run() == 5


def test_filter_result_test(ctx, ops):  # :33-61
    keep = [5, 8, 9, 100, 270]
    v = (np.arange(4096, dtype=np.uint64) + (1 << 30)).astype(np.uint32)
    v[keep] = keep
    f = ops.FilterGpu(ctx, [v])
    f.Prepare()
    chunks = f.GetResult()
    assert len(chunks) == 1 and chunks[0].tolist() == keep


def test_filter_longer_test(ctx, ops):  # :63-78
    batches = gen_batches(1, 65536)
    f = ops.FilterGpu(ctx, batches)
    f.Prepare()
    chunks = f.GetResult()
    assert np.array_equal(chunks[0], oracle.filter_lt(batches[0]))
    assert chunks[0].size == 16358  # SURVEY.md §8(c) fingerprint of RandomArrayGenerator(42)
    t = f.Timers()
    assert t is not None


def _filter_into(ctx, batches, thr=1 << 30, capacity=None):
    from dpu_olap_b200._lib import Timings
    nb = len(batches)
    ptrs = (C.c_void_p * max(nb, 1))(*[b.ctypes.data for b in batches])
    lens = (C.c_int64 * max(nb, 1))(*[b.size for b in batches])
    n = sum(b.size for b in batches)
    cap = n if capacity is None else capacity
    out = np.full(max(cap, 1), 0xDEADBEEF, dtype=np.uint32)
    counts = (C.c_int64 * max(nb, 1))()
    total = C.c_uint64(0)
    t = Timings()
    rc = ctx._lib.b2_filter_lt_u32_host_into(ctx._h, ptrs, lens, nb, thr, out.ctypes.data, cap, counts,
                                             C.byref(total), C.byref(t))
    return rc, out, [counts[b] for b in range(nb)], int(total.value), t


@pytest.mark.parametrize("nb,bl", [(0, 0), (1, 5), (3, 4096), (64, 65536), (600, 65536), (5, 1 << 22)])
def test_filter_host_into_uniform(ctx, nb, bl):
    rng = np.random.default_rng(nb * 7 + bl)
    batches = [rng.integers(0, 2**32, size=bl, dtype=np.uint32) for _ in range(nb)]
    rc, out, counts, total, t = _filter_into(ctx, batches)
    assert rc == 0
    exp = [oracle.filter_lt(b) for b in batches]
    assert counts == [e.size for e in exp]
    assert total == sum(counts)
    if nb:
        assert np.array_equal(out[:total], np.concatenate(exp) if total else np.empty(0, np.uint32))
        assert t.h2d_bytes == 4 * nb * bl


def test_filter_host_into_ragged_many_chunks(ctx):
    rng = np.random.default_rng(5)
    lens = [0, 1, 70000, 3, 0, (1 << 24) + 5, 65536, 12345, (1 << 24) + 1, 2]
    batches = [rng.integers(0, 2**32, size=n, dtype=np.uint32) for n in lens]
    for thr in (1 << 30, 0, 0xFFFFFFFF):
        rc, out, counts, total, _ = _filter_into(ctx, batches, thr)
        assert rc == 0
        exp = [oracle.filter_lt(b, thr) for b in batches]
        assert counts == [e.size for e in exp]
        assert np.array_equal(out[:total], np.concatenate(exp))


def test_filter_host_into_overflow_is_reported(ctx):
    batches = gen_batches(4, 65536)
    rc, out, counts, total, _ = _filter_into(ctx, batches, capacity=100)
    assert rc == 6  # B2_ERR_OVERFLOW
    assert total == sum(oracle.filter_lt(b).size for b in batches)
    assert b"output" in ctx._lib.b2_last_error(ctx._h)


def test_filter_two_step_api_matches_streaming(ctx, ops):
    batches = gen_batches(40, 65536)
    f = ops.FilterGpu(ctx, batches)
    f.Prepare()
    chunks = f.GetResult()
    rc, out, counts, total, _ = _filter_into(ctx, batches)
    assert rc == 0 and [c.size for c in chunks] == counts
    assert np.array_equal(np.concatenate(chunks), out[:total])


# ---- SumTest (host/aggr/aggr_test.cc) ---------------------------------------------------------------
def test_sum_simple_test(ctx, ops):  # :24-36
    s = ops.SumGpu(ctx, [np.array([0, 2, 3, 8, 9], np.uint32)])
    s.Prepare()
    assert s.Run() == 22


def test_sum_large_test(ctx, ops):  # :38-49 (128 x 65536)
    batches = gen_batches(128, 65536)
    s = ops.SumGpu(ctx, batches)
    s.Prepare()
    assert s.Run() == sum(oracle.sum_u32(b) for b in batches)


# ---- TakeTest (host/take/take_test.cc) --------------------------------------------------------------
def test_take_simple_test(ctx, ops):  # :24-36
    t = ops.TakeGpu(ctx, [np.array([0, 2, 3, 8, 9], np.uint32)], [np.array([0, 1, 4], np.uint32)])
    t.Prepare()
    assert t.Run()[0].tolist() == [0, 2, 9]


def test_take_large_test(ctx, ops):  # :38-72 (128 x 65536 values, 128 x 8192 indices in [0, 65535])
    g = oracle.RandomArrayGenerator(42)
    vals = oracle.make_random_batches(g, 128, 65536)
    idx = oracle.make_random_batches(g, 128, 8192, 0, 65535)
    t = ops.TakeGpu(ctx, vals, idx)
    t.Prepare()
    out = t.Run()
    for b in range(128):
        assert np.array_equal(out[b], oracle.take(vals[b], idx[b]))


# ---- JoinTest (host/join/join_test.cc) --------------------------------------------------------------
def test_join_simple_test(ctx, ops):  # :40-80: two 5-row batches per side
    left = [{"fk": np.array([0, 2, 3, 8, 9], np.uint32), "v_l": np.array([100, 102, 103, 108, 109], np.uint32)},
            {"fk": np.array([10, 12, 13, 18, 19], np.uint32), "v_l": np.array([110, 112, 113, 118, 119], np.uint32)}]
    right = [{"pk": np.array([3, 8, 9, 0, 2], np.uint32), "v_r": np.array([53, 58, 59, 50, 52], np.uint32)},
             {"pk": np.array([12, 13, 18, 19, 10], np.uint32), "v_r": np.array([62, 63, 68, 69, 60], np.uint32)}]
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    out = j.Run()
    rows = sorted(zip(out["fk"].tolist(), out["v_l"].tolist(), out["v_r"].tolist()))
    assert rows == [(0, 100, 50), (2, 102, 52), (3, 103, 53), (8, 108, 58), (9, 109, 59), (10, 110, 60),
                    (12, 112, 62), (13, 113, 63), (18, 118, 68), (19, 119, 69)]


def test_join_large_test(ctx, ops):  # :82-121, at 32 x 65536 so the oracle sorts in a second
    nb, bs = 32, 65536
    g = oracle.RandomArrayGenerator(42)
    x = oracle.make_random_batches(g, nb, bs)
    y = oracle.make_random_batches(g, nb, bs)
    fk = oracle.make_fk_batches(g, bs, nb, bs)
    pk = oracle.make_index_batches(nb, bs)
    left = [{"fk": fk[b], "y": y[b]} for b in range(nb)]
    right = [{"pk": pk[b], "x": x[b]} for b in range(nb)]
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    out = j.Run()
    assert out["fk"].size == nb * bs
    exp = oracle.sort_rows(*oracle.join(np.concatenate(fk), np.concatenate(y), np.concatenate(pk),
                                        np.concatenate(x)))
    got = oracle.sort_rows(out["fk"], out["y"], out["x"])
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)


# ---- PartitionTest (host/partition/partition_test.cc; skipped in the reference) ----------------------
def test_partition_simple_test(ctx, ops):  # :21-57: {0,2,3,8} -> 3 / 1 rows, sums 13 / 406
    batches = [{"pk": np.array([0, 2], np.uint32), "x": np.array([100, 101], np.uint32)},
               {"pk": np.array([3, 8], np.uint32), "x": np.array([102, 103], np.uint32)}]
    p = ops.PartitionGpu(ctx, batches, 2, "pk")
    p.Prepare()
    parts = p.Run()
    assert sorted(len(q["pk"]) for q in parts) == [1, 3]
    assert sum(int(q["pk"].sum()) for q in parts) == 13
    assert sum(int(q["x"].sum()) for q in parts) == 406
    for pid, q in enumerate(parts):
        assert all(oracle.bucket(int(k), 2) == pid for k in q["pk"])


def test_partition_large_test(ctx, ops):  # :59-92: 32 partitions within 10 % of the mean
    nb, bs = 128, 65536
    g = oracle.RandomArrayGenerator(42)
    x = oracle.make_random_batches(g, nb, bs)
    pk = oracle.make_index_batches(nb, bs)
    p = ops.PartitionGpu(ctx, [{"pk": pk[b], "x": x[b]} for b in range(nb)], 32, "pk")
    p.Prepare()
    parts = p.Run()
    mean = nb * bs / 32
    assert sum(q["pk"].size for q in parts) == nb * bs
    allx = np.concatenate(x)
    for pid, q in enumerate(parts):
        assert abs(q["pk"].size - mean) < 0.1 * mean
        assert np.array_equal(oracle.partition_ids(q["pk"], 32), np.full(q["pk"].size, pid))
        assert np.array_equal(allx[q["pk"]], q["x"])  # columns stay aligned (pk is the row index)


def test_filter_host_into_pinned_inputs_use_the_gather_upload(ctx):
    """Separately allocated, page-locked batches: with b2_ctx_set_inputs_pinned one gather kernel per
    group reads them over PCIe; the result is the same as with per-batch DMA copies."""
    rng = np.random.default_rng(99)
    batches = [rng.integers(0, 2**32, size=65536 if b % 7 else 12345, dtype=np.uint32) for b in range(200)]
    exp = [oracle.filter_lt(b) for b in batches]
    lib = ctx._lib
    for b in batches:
        assert lib.b2_host_register(b.ctypes.data, b.nbytes) == 0
    try:
        l0 = ctx.launches
        rc, out, counts, total, t0 = _filter_into(ctx, batches)
        plain_launches = ctx.launches - l0
        assert rc == 0 and counts == [e.size for e in exp]
        assert lib.b2_ctx_set_inputs_pinned(ctx._h, 1) == 0
        l0 = ctx.launches
        rc, out2, counts2, total2, t1 = _filter_into(ctx, batches)
        assert rc == 0 and counts2 == counts and total2 == total
        assert np.array_equal(out2[:total], np.concatenate(exp))
        assert ctx.launches - l0 > plain_launches          # the gather kernels were launched
        assert t1.h2d_bytes == t0.h2d_bytes
    finally:
        lib.b2_ctx_set_inputs_pinned(ctx._h, 0)
        for b in batches:
            lib.b2_host_unregister(b.ctypes.data)


def test_join_run_aggregate_fused_pipeline(ctx, ops):
    """JoinGpu.RunAggregate == aggregates of JoinGpu.Run()'s columns (and of the oracle's join), with
    and without the pushed-down filter on the probe side's payload."""
    g = oracle.RandomArrayGenerator(42)
    nb, bs = 8, 1 << 16
    x = oracle.make_random_batches(g, nb, bs)
    y = oracle.make_random_batches(g, nb, bs)
    fk = oracle.make_fk_batches(g, bs, nb, bs)
    pk = oracle.make_index_batches(nb, bs)
    left = [{"fk": fk[b], "y": y[b]} for b in range(nb)]
    right = [{"pk": pk[b], "x": x[b]} for b in range(nb)]
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    cols = j.Run()
    agg = j.RunAggregate()
    assert agg == {"rows": cols["fk"].size, "sum_y": int(cols["y"].astype(np.uint64).sum(dtype=np.uint64)),
                   "sum_x": int(cols["x"].astype(np.uint64).sum(dtype=np.uint64))}
    assert j.Timers()["copy-from-dpu"] == 0.0 or j._last[0].d2h_bytes == 24
    thr = 1 << 30
    exp = oracle.join_aggr(np.concatenate(fk), np.concatenate(y), np.concatenate(pk), np.concatenate(x), thr)
    assert j.RunAggregate(thr) == exp
    assert 0 < exp["rows"] < cols["fk"].size


# ---- JoinDpu takes every non-key column of both sides (join_dpu.cc:127-138,325-341) ------------------
@pytest.mark.parametrize("nlp,nrp", [(2, 1), (1, 3), (3, 2), (0, 1), (2, 0)])
def test_join_many_payload_columns(ctx, ops, nlp, nrp):
    rng = np.random.default_rng(10 * nlp + nrp)
    nb, bs = 6, 30_000
    n = nb * bs
    pk = rng.permutation(np.arange(n, dtype=np.uint32) * 3)
    pk[:1000] = pk[1000:2000]                                   # duplicate build keys
    fk = rng.integers(0, 3 * n + 50, size=n, dtype=np.uint32)   # a third of the probe rows match
    lcols = {f"l{c}": rng.integers(0, 2**32, size=n, dtype=np.uint32) for c in range(nlp)}
    rcols = {f"r{c}": rng.integers(0, 2**32, size=n, dtype=np.uint32) for c in range(nrp)}
    sl = lambda a, b: a[b * bs:(b + 1) * bs]
    left = [{"fk": sl(fk, b), **{k: sl(v, b) for k, v in lcols.items()}} for b in range(nb)]
    right = [{"pk": sl(pk, b), **{k: sl(v, b) for k, v in rcols.items()}} for b in range(nb)]
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    out = j.Run()
    assert list(out) == ["fk"] + list(lcols) + list(rcols)
    # expected: the oracle joins row numbers; every payload column follows its side's row number
    ids = np.arange(n, dtype=np.uint32)
    e_fk, e_l, e_r = oracle.join(fk, ids, pk, ids)
    exp = [e_fk] + [v[e_l] for v in lcols.values()] + [v[e_r] for v in rcols.values()]
    got = [out[k] for k in out]
    assert got[0].size == e_fk.size
    order_g = np.lexsort(tuple(reversed(got)))
    order_e = np.lexsort(tuple(reversed(exp)))
    for g, e in zip(got, exp):
        assert np.array_equal(g[order_g], e[order_e])


# ---- 64-bit keys (HT_64BIT_KEYS, hashtable.h:14-18) and wide / typed payload columns -----------------
def _fold(k):
    k = np.asarray(k, dtype=np.uint64)
    return ((k & np.uint64(0xFFFFFFFF)) ^ (((k >> np.uint64(32)) * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF))).astype(np.uint32)


@pytest.mark.parametrize("key_dtype", [np.uint64, np.int64, np.int32])
def test_join_wide_keys_and_typed_payloads_match_arrow(ctx, ops, key_dtype):
    import pyarrow as pa
    rng = np.random.default_rng(int(np.dtype(key_dtype).itemsize))
    nb, bs = 4, 50_000
    n = nb * bs
    if np.dtype(key_dtype).itemsize == 8:
        pk = rng.integers(0, 2**63, size=n, dtype=np.uint64)
        pk[1] = pk[0]                                             # a duplicate build key
        # build keys with the SAME 32-bit fold as other build keys but different values: the probe finds them
        # as candidates and the 64-bit verification must reject them
        hi = rng.integers(1, 2**31, size=2000, dtype=np.uint64)
        lo = (_fold(pk[2:2002]).astype(np.uint64) ^ ((hi * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)))
        pk[2002:4002] = (hi << np.uint64(32)) | lo
        assert np.array_equal(_fold(pk[2002:4002]), _fold(pk[2:2002]))
        fk = pk[rng.integers(0, n, size=n)]
        fk[:500] = rng.integers(0, 2**63, size=500, dtype=np.uint64)  # no match
        pk, fk = pk.astype(key_dtype), fk.astype(key_dtype)
    else:
        pk = (rng.permutation(n) - n // 2).astype(key_dtype)       # negative int32 keys too
        fk = pk[rng.integers(0, n, size=n)]
    y64 = rng.integers(-2**62, 2**62, size=n, dtype=np.int64)
    yf = rng.random(n).astype(np.float32)
    x32 = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    xd = rng.random(n)                                            # float64
    sl = lambda a, b: a[b * bs:(b + 1) * bs]
    left = [pa.record_batch([pa.array(sl(fk, b)), pa.array(sl(y64, b)), pa.array(sl(yf, b))], names=["fk", "y64", "yf"])
            for b in range(nb)]
    right = [pa.record_batch([pa.array(sl(pk, b)), pa.array(sl(x32, b)), pa.array(sl(xd, b))], names=["pk", "x32", "xd"])
             for b in range(nb)]
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    out = j.Run()
    assert list(out) == ["fk", "y64", "yf", "x32", "xd"]
    assert out["fk"].dtype == np.dtype(key_dtype) and out["y64"].dtype == np.int64 and out["xd"].dtype == np.float64
    exp = pa.Table.from_batches(left).join(pa.Table.from_batches(right), keys="fk", right_keys="pk", join_type="inner")
    assert out["fk"].size == exp.num_rows
    want = [exp.column(c).combine_chunks().to_numpy(zero_copy_only=False) for c in out]
    got = [out[c] for c in out]
    og, ow = np.lexsort(tuple(reversed(got))), np.lexsort(tuple(reversed(want)))
    for g, w in zip(got, want):
        assert np.array_equal(g[og], w[ow])
