"""GPU parity tests of the nullable entry points against the CPU oracle, which
tests/test_oracle_nullable.py pins to Arrow's kernels. Bit-exact: filter output and chunk
boundaries, all four aggregates, take values AND result bitmap."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
pa = pytest.importorskip("pyarrow")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def dev_bits(valid):
    return torch.from_numpy(oracle.pack_bits(valid)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def make(rng, n, null_frac, hi=2**32):
    return rng.integers(0, hi, size=n, dtype=np.uint32), rng.random(n) >= null_frac


# ---- filter --------------------------------------------------------------------------------------
@pytest.mark.parametrize("nb,bl,null_frac,thr", [
    (1, 1, 0.0, 1 << 30), (1, 1, 1.0, 1 << 30), (3, 4096, 0.3, 1 << 30), (16, 65536, 0.1, 1 << 30),
    (5, 4100, 0.5, 1 << 31),        # batch length not a multiple of 32: tiles start mid-word
    (7, 8192 + 36, 0.25, 1 << 30),  # multiple of 4 but not of 32: TMA tiles on the row-by-row bitmap path
    (4, 12288, 0.9, 0xFFFFFFFF), (4, 12288, 0.5, 0), (64, 65536, 0.01, 1 << 29), (2, 100_003, 0.5, 1 << 30)])
def test_filter_nullable_dev(ctx, nb, bl, null_frac, thr):
    rng = np.random.default_rng(nb * 131 + bl)
    v, valid = make(rng, nb * bl, null_frac)
    out, end, total = ctx.filter_nullable_dev(dev(v), dev_bits(valid), nb, bl, thr)
    torch.cuda.synchronize()
    exp = [oracle.filter_lt_nullable(v[b * bl:(b + 1) * bl], valid[b * bl:(b + 1) * bl], thr) for b in range(nb)]
    assert np.array_equal(end.cpu().numpy()[:nb], np.cumsum([e.size for e in exp]))
    n = int(total.cpu()[0])
    assert n == sum(e.size for e in exp)
    assert np.array_equal(host(out)[:n], np.concatenate(exp))


def test_filter_nullable_none_bitmap_equals_plain(ctx):
    rng = np.random.default_rng(5)
    v = rng.integers(0, 2**32, size=8 * 65536, dtype=np.uint32)
    a, _, ta = ctx.filter_nullable_dev(dev(v), None, 8, 65536, 1 << 30)
    b, _, tb = ctx.filter_dev(dev(v), 8, 65536, 1 << 30)
    torch.cuda.synchronize()
    n = int(ta.cpu()[0])
    assert n == int(tb.cpu()[0]) and np.array_equal(host(a)[:n], host(b)[:n])


# ---- aggregates ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,null_frac,misalign", [(0, 0.0, 0), (1, 0.0, 0), (5, 1.0, 0), (1000, 0.3, 0),
                                                    (1 << 20, 0.5, 0), ((1 << 22) + 13, 0.05, 0),
                                                    (100_000, 0.5, 1), (100_000, 0.0, 3), (70_000, 0.999, 0)])
def test_aggregates_dev(ctx, n, null_frac, misalign):
    from dpu_olap_b200.ops import decode_aggr
    rng = np.random.default_rng(n + misalign)
    v, valid = make(rng, n, null_frac)
    buf = dev(np.concatenate([np.zeros(misalign, np.uint32), v]))  # values not 16 B aligned when misalign > 0
    col = buf[misalign:]
    got = decode_aggr(ctx.aggr_dev(col, dev_bits(valid)))
    assert got == oracle.aggr_nullable(v, valid)
    got = decode_aggr(ctx.aggr_dev(col, None))  # no bitmap: every row counts
    assert got == oracle.aggr_nullable(v, np.ones(n, bool))
    if n:
        assert got["sum"] == oracle.sum_u32(v)


# ---- take --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nb,vl,il,fv,fi", [(1, 1, 1, 0.0, 0.0), (3, 1000, 37, 0.3, 0.3), (8, 65536, 8192, 0.5, 0.1),
                                             (4, 4096, 10_000, 0.0, 0.5), (4, 4096, 10_000, 0.5, 0.0),
                                             (2, 100, 100, 1.0, 0.0), (2, 100, 100, 0.0, 1.0),
                                             (16, 1 << 18, 1 << 15, 0.2, 0.2)])
def test_take_nullable_dev(ctx, nb, vl, il, fv, fi):
    rng = np.random.default_rng(nb + vl + il)
    v, vvalid = make(rng, nb * vl, fv)
    i, ivalid = make(rng, nb * il, fi, hi=vl)
    out, obits = ctx.take_nullable_dev(dev(v), dev_bits(vvalid), vl, dev(i), dev_bits(ivalid), il, nb)
    torch.cuda.synchronize()
    got, gbits = host(out), oracle.unpack_bits(obits.cpu().numpy(), nb * il)
    for b in range(nb):
        e_out, e_ok = oracle.take_nullable(v[b * vl:(b + 1) * vl], vvalid[b * vl:(b + 1) * vl],
                                           i[b * il:(b + 1) * il], ivalid[b * il:(b + 1) * il])
        assert np.array_equal(gbits[b * il:(b + 1) * il], e_ok)
        assert np.array_equal(got[b * il:(b + 1) * il], e_out)
    # one side without a bitmap
    out2, obits2 = ctx.take_nullable_dev(dev(v), None, vl, dev(i), dev_bits(ivalid), il, nb)
    torch.cuda.synchronize()
    assert np.array_equal(oracle.unpack_bits(obits2.cpu().numpy(), nb * il), ivalid)


# ---- operator classes over Arrow arrays with nulls (host buffers through the C ABI) -------------------
def arrow_batches(rng, nb, bl, null_frac, hi=2**32, slice_off=0):
    """pyarrow uint32 arrays with nulls; slice_off > 0 makes them views with a bitmap BIT offset."""
    out = []
    for _ in range(nb):
        v, valid = make(rng, bl + slice_off, null_frac, hi)
        out.append(pa.array(v, type=pa.uint32(), mask=~valid).slice(slice_off, bl))
    return out


def np_of(arr):
    return (arr.fill_null(0).to_numpy(zero_copy_only=False).astype(np.uint32), ~np.asarray(arr.is_null()))


@pytest.mark.parametrize("nb,bl,slice_off", [(4, 65536, 0), (3, 4096, 5), (6, 8192 + 36, 0), (2, 1000, 3)])
def test_filter_gpu_with_arrow_nulls(ctx, nb, bl, slice_off):
    import pyarrow.compute as pc
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(nb * bl + slice_off)
    batches = arrow_batches(rng, nb, bl, 0.3, slice_off=slice_off)
    f = ops.FilterGpu(ctx, batches)
    f.Prepare()
    chunks = f.GetResult()
    for b, arr in enumerate(batches):
        v, valid = np_of(arr)
        assert np.array_equal(chunks[b], oracle.filter_lt_nullable(v, valid))
        # and directly against Arrow on the same array (the reference's FilterNative plan)
        exp = pc.filter(arr, pc.less(arr, pa.scalar(1 << 30, pa.uint32())))
        assert np.array_equal(chunks[b], exp.to_numpy(zero_copy_only=False).astype(np.uint32))
    assert f.Run() == sum(c.size for c in chunks)
    assert f.Timers()["total"] > 0


def test_sum_gpu_with_arrow_nulls_and_aggregates(ctx):
    import pyarrow.compute as pc
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(77)
    batches = arrow_batches(rng, 5, 50_000, 0.4, slice_off=7)
    s = ops.SumGpu(ctx, batches)
    s.Prepare()
    whole = pa.chunked_array(batches)
    assert s.Run() == pc.sum(whole.cast(pa.uint64())).as_py()
    agg = s.Aggregates()
    mm = pc.min_max(whole)
    assert agg == {"sum": pc.sum(whole.cast(pa.uint64())).as_py(), "count": pc.count(whole).as_py(),
                   "min": mm["min"].as_py(), "max": mm["max"].as_py()}
    all_null = ops.SumGpu(ctx, [pa.array([None, None, None], type=pa.uint32())])
    assert all_null.Run() is None and all_null.Aggregates()["count"] == 0  # Arrow: sum of no rows is null
    masked = np.ma.MaskedArray(np.arange(10, dtype=np.uint32), mask=[0, 1] * 5)
    assert ops.SumGpu(ctx, [masked]).Run() == 0 + 2 + 4 + 6 + 8


def test_take_gpu_with_arrow_nulls(ctx):
    import pyarrow.compute as pc
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(78)
    vals = arrow_batches(rng, 4, 4096, 0.3, slice_off=2)
    idx = arrow_batches(rng, 4, 1000, 0.2, hi=4096, slice_off=1)
    t = ops.TakeGpu(ctx, vals, idx)
    t.Prepare()
    res = t.Run()
    for b in range(4):
        exp = pc.take(vals[b], idx[b])  # TakeNative's call, take_native.cc:27
        assert res[b].type == pa.uint32() and res[b].null_count == exp.null_count
        assert res[b].equals(exp)
    # nulls on one side only, and the all-valid call still returns plain numpy arrays
    plain_v = [np.asarray(a.fill_null(0)) for a in vals]
    res2 = ops.TakeGpu(ctx, plain_v, idx).Run()
    for b in range(4):
        assert res2[b].equals(pc.take(pa.array(plain_v[b], type=pa.uint32()), idx[b]))
    res3 = ops.TakeGpu(ctx, plain_v, [np.asarray(a.fill_null(0)) for a in idx]).Run()
    assert isinstance(res3[0], np.ndarray)


def test_nullable_ragged_batches_and_loud_rejections(ctx):
    import pyarrow.compute as pc
    from dpu_olap_b200 import ops
    from dpu_olap_b200._lib import B2Error
    rng = np.random.default_rng(17)
    # batches of different lengths (one empty) with nulls: the ragged kernel path with a bitmap
    batches = []
    for n in (3, 0, 5000, 4096, 70_001, 1):
        v, valid = make(rng, n, 0.3)
        batches.append(pa.array(v, type=pa.uint32(), mask=~valid))
    chunks = ops.FilterGpu(ctx, batches).GetResult()
    for arr, got in zip(batches, chunks):
        exp = pc.filter(arr, pc.less(arr, pa.scalar(1 << 30, pa.uint32())))
        assert np.array_equal(got, exp.to_numpy(zero_copy_only=False).astype(np.uint32))
    ragged_take = [pa.array([1, None, 3], type=pa.uint32()), pa.array([1, None], type=pa.uint32())]
    with pytest.raises(B2Error) as e:   # the nullable take still wants equal batch lengths
        ops.TakeGpu(ctx, ragged_take, [pa.array([0, None, 1], type=pa.uint32()), pa.array([0, 1], type=pa.uint32())]).Run()
    assert e.value.status == 4  # B2_ERR_UNSUPPORTED
    batches = ragged_take
    with pytest.raises(ValueError):  # nullable PAYLOAD columns have no semantics in the join (keys do, below)
        ops.JoinGpu(ctx, [{"fk": batches[0], "y": batches[0]}], [{"pk": batches[0], "x": batches[0]}])
    plain = pa.array([7, 8, 9], type=pa.uint32())
    out = ops.JoinGpu(ctx, [{"fk": batches[0], "y": plain}], [{"pk": batches[0], "x": plain}]).Run()
    assert sorted(zip(out["fk"].tolist(), out["y"].tolist(), out["x"].tolist())) == [(1, 7, 7), (3, 9, 9)]  # null != null


# ---- other 32-bit types (b2_filter_lt_32_dev) ----------------------------------------------------------
def typed_column(rng, dtype, n):
    if dtype == np.float32:
        v = (rng.standard_normal(n) * 3).astype(np.float32)
        v[::97] = np.nan
        v[::101] = np.inf
        v[::103] = -np.inf
        v[::107] = -0.0
        return v
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)


@pytest.mark.parametrize("dtype,thr", [(np.int32, -5), (np.int32, 0), (np.int32, 2**31 - 1), (np.int32, -2**31),
                                        (np.float32, 0.0), (np.float32, -1.5), (np.float32, np.inf),
                                        (np.float32, np.nan), (np.uint32, 1 << 30)])
@pytest.mark.parametrize("nb,bl,null_frac", [(6, 65536, 0.0), (5, 4100, 0.3), (3, 12288, 0.5)])
def test_filter_typed_dev(ctx, dtype, thr, nb, bl, null_frac):
    rng = np.random.default_rng(nb * bl + int(null_frac * 10))
    v = typed_column(rng, dtype, nb * bl)
    valid = rng.random(nb * bl) >= null_frac
    t = torch.from_numpy(v.view(np.int32)).cuda()
    out, end, total = ctx.filter_typed_dev(t, dtype, thr, nb, bl, valid=dev_bits(valid) if null_frac else None)
    torch.cuda.synchronize()
    exp = [oracle.filter_lt_typed(v[b * bl:(b + 1) * bl], valid[b * bl:(b + 1) * bl], thr) for b in range(nb)]
    n = int(total.cpu()[0])
    assert n == sum(e.size for e in exp)
    assert np.array_equal(end.cpu().numpy()[:nb], np.cumsum([e.size for e in exp]))
    assert np.array_equal(host(out)[:n], np.concatenate(exp).view(np.uint32))  # bit-exact, NaN payloads included


@pytest.mark.parametrize("pa_type,thr", [("int32", -1000), ("float32", 0.25)])
def test_filter_gpu_other_types_against_arrow(ctx, pa_type, thr):
    import pyarrow.compute as pc
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(31)
    dtype = np.dtype(pa_type)
    batches = []
    for _ in range(4):
        v = typed_column(rng, dtype.type, 8192 + 3)
        batches.append(pa.array(v, mask=rng.random(v.size) < 0.25).slice(3, 8192))
    f = ops.FilterGpu(ctx, batches, threshold=thr)
    f.Prepare()
    chunks = f.GetResult()
    for b, arr in enumerate(batches):
        exp = pc.filter(arr, pc.less(arr, pa.scalar(thr, arr.type))).to_numpy(zero_copy_only=False).astype(dtype)
        assert chunks[b].dtype == dtype and np.array_equal(chunks[b].view(np.uint32), exp.view(np.uint32))
    assert f.Run() == sum(c.size for c in chunks)
    with pytest.raises(TypeError):
        ops.SumGpu(ctx, [np.zeros(4, np.float32)])   # float aggregates are order dependent: rejected
    with pytest.raises(TypeError):
        ops.FilterGpu(ctx, [np.zeros(4, np.int32), np.zeros(4, np.float32)])


@pytest.mark.parametrize("n,null_frac", [(1, 0.0), (5, 1.0), (100_000, 0.3), ((1 << 21) + 7, 0.0)])
def test_int32_aggregates(ctx, n, null_frac):
    import pyarrow.compute as pc
    from dpu_olap_b200 import ops
    from dpu_olap_b200.ops import decode_aggr
    rng = np.random.default_rng(n)
    v = rng.integers(-2**31, 2**31 - 1, size=n, dtype=np.int32, endpoint=True)
    valid = rng.random(n) >= null_frac
    t = torch.from_numpy(v).cuda()
    got = decode_aggr(ctx.aggr_dev(t, dev_bits(valid) if null_frac else None, dtype=np.int32), np.int32)
    assert got == oracle.aggr_nullable(v, valid if null_frac else np.ones(n, bool))
    # the operator class over Arrow int32 batches, against Arrow itself
    arr = pa.array(v, type=pa.int32(), mask=~valid)
    s = ops.SumGpu(ctx, [arr])
    mm = pc.min_max(arr)
    assert s.Aggregates() == {"sum": pc.sum(arr).as_py(), "count": pc.count(arr).as_py(),
                              "min": mm["min"].as_py(), "max": mm["max"].as_py()}
    assert s.Run() == pc.sum(arr).as_py()


# ---- join with nullable keys: Arrow's inner hash join never matches a null key (join_native.cc:31-36) ----
@pytest.mark.parametrize("null_left,null_right", [(True, False), (False, True), (True, True)])
def test_join_nullable_keys_match_arrow(ctx, null_left, null_right):
    import pyarrow as pa

    from dpu_olap_b200 import ops
    rng = np.random.default_rng(3 + 2 * null_left + null_right)
    nb, bs = 5, 40_000
    n = nb * bs
    pk = rng.permutation(n).astype(np.uint32)
    fk = rng.integers(0, n + n // 4, size=n, dtype=np.uint32)   # some probe keys have no match
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    lmask = rng.random(n) < 0.1 if null_left else np.zeros(n, bool)
    rmask = rng.random(n) < 0.1 if null_right else np.zeros(n, bool)

    def batches(key_name, key, mask, pay_name, pay, slice_off):
        out = []
        for b in range(nb):
            sl = slice(b * bs, (b + 1) * bs)
            # a sliced array: the validity bitmap starts at a bit offset
            k = pa.array(np.concatenate([np.zeros(slice_off, np.uint32), key[sl]]),
                         mask=np.concatenate([np.zeros(slice_off, bool), mask[sl]])).slice(slice_off)
            out.append(pa.record_batch([k, pa.array(pay[sl])], names=[key_name, pay_name]))
        return out
    left = batches("fk", fk, lmask, "y", y, 3)
    right = batches("pk", pk, rmask, "x", x, 5)
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    out = j.Run()
    lt, rt = pa.Table.from_batches(left), pa.Table.from_batches(right)
    exp = lt.join(rt, keys="fk", right_keys="pk", join_type="inner")
    assert out["fk"].size == exp.num_rows
    got = oracle.sort_rows(out["fk"], out["y"], out["x"])
    want = oracle.sort_rows(*[exp.column(c).combine_chunks().to_numpy(zero_copy_only=False).astype(np.uint32)
                              for c in ("fk", "y", "x")])
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # and against the oracle on the rows whose keys are valid
    e2 = oracle.sort_rows(*oracle.join(fk[~lmask], y[~lmask], pk[~rmask], x[~rmask]))
    for a, b in zip(got, e2):
        assert np.array_equal(a, b)
