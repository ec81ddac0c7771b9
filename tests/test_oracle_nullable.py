"""Pins the nullable restatements in oracle/oracle.py against Arrow's own kernels (pyarrow.compute,
Arrow 24 — the library the reference's Native classes run their plans on: filter_native.cc:52-66,
aggr_native.cc:68-73, take_native.cc:27). CPU only."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import oracle


def make(rng, n, null_frac, hi=2**32):
    v = rng.integers(0, hi, size=n, dtype=np.uint32)
    valid = rng.random(n) >= null_frac
    return v, valid, pa.array(v, type=pa.uint32(), mask=~valid)


@pytest.mark.parametrize("n,null_frac", [(0, 0.0), (1, 1.0), (1, 0.0), (1000, 0.1), (4097, 0.5), (70000, 0.9),
                                         (5000, 1.0), (5000, 0.0)])
def test_filter_matches_arrow(n, null_frac):
    rng = np.random.default_rng(n + int(null_frac * 100))
    v, valid, arr = make(rng, n, null_frac)
    for thr in (1 << 30, 0, 1, 0xFFFFFFFF):
        exp = pc.filter(arr, pc.less(arr, pa.scalar(thr, pa.uint32())))  # null predicate drops the row
        assert exp.null_count == 0
        got = oracle.filter_lt_nullable(v, valid, thr)
        assert np.array_equal(got, exp.to_numpy(zero_copy_only=False).astype(np.uint32))


@pytest.mark.parametrize("n,null_frac", [(0, 0.0), (7, 1.0), (1, 0.0), (1000, 0.1), (65536, 0.5), (70001, 0.99)])
def test_aggregates_match_arrow(n, null_frac):
    rng = np.random.default_rng(100 + n)
    v, valid, arr = make(rng, n, null_frac)
    got = oracle.aggr_nullable(v, valid)
    assert got["sum"] == pc.sum(arr.cast(pa.uint64()) if n else arr).as_py()
    assert got["count"] == pc.count(arr).as_py()
    mm = pc.min_max(arr)
    assert got["min"] == mm["min"].as_py() and got["max"] == mm["max"].as_py()


@pytest.mark.parametrize("nv,ni,fv,fi", [(1, 0, 0.0, 0.0), (10, 30, 0.3, 0.3), (4096, 1000, 0.0, 0.5),
                                          (4096, 5000, 0.5, 0.0), (100, 100, 1.0, 0.0), (100, 100, 0.0, 1.0)])
def test_take_matches_arrow(nv, ni, fv, fi):
    rng = np.random.default_rng(nv * 3 + ni)
    v, vvalid, varr = make(rng, nv, fv)
    i, ivalid, iarr = make(rng, ni, fi, hi=nv)
    exp = pc.take(varr, iarr)
    out, ok = oracle.take_nullable(v, vvalid, i, ivalid)
    assert np.array_equal(ok, ~np.asarray(exp.is_null()))
    assert np.array_equal(out[ok], exp.drop_null().to_numpy().astype(np.uint32))
    assert not out[~ok].any()


def test_bitmap_round_trip_and_arrow_layout():
    rng = np.random.default_rng(9)
    valid = rng.random(1003) > 0.4
    bits = oracle.pack_bits(valid)
    assert bits.size % 4 == 0 and np.array_equal(oracle.unpack_bits(bits, valid.size), valid)
    arr = pa.array(np.arange(1003, dtype=np.uint32), mask=~valid)
    arrow_bits = np.frombuffer(arr.buffers()[0], dtype=np.uint8)
    assert np.array_equal(oracle.unpack_bits(arrow_bits, valid.size), valid)  # same bit order as Arrow
    sl = arr.slice(5, 900)  # a sliced array keeps the buffer and carries a bit offset
    assert np.array_equal(oracle.unpack_bits(np.frombuffer(sl.buffers()[0], dtype=np.uint8), 900, sl.offset),
                          valid[5:905])


@pytest.mark.parametrize("dtype,thr", [(np.int32, -5), (np.int32, 0), (np.int32, 2**31 - 1), (np.int32, -2**31),
                                        (np.float32, 0.0), (np.float32, -1.5), (np.float32, np.inf),
                                        (np.float32, np.nan), (np.uint32, 1 << 30)])
def test_typed_filter_matches_arrow(dtype, thr):
    rng = np.random.default_rng(5)
    n = 20_000
    if dtype == np.float32:
        v = rng.standard_normal(n).astype(np.float32) * 3
        v[::97] = np.nan
        v[::101] = np.inf
        v[::103] = -np.inf
        v[::107] = -0.0
    else:
        info = np.iinfo(dtype)
        v = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    valid = rng.random(n) > 0.2
    arr = pa.array(v, mask=~valid)
    exp = pc.filter(arr, pc.less(arr, pa.scalar(thr, arr.type)))
    got = oracle.filter_lt_typed(v, valid, thr)
    assert got.dtype == v.dtype and np.array_equal(got.view(np.uint32),
                                                   exp.to_numpy(zero_copy_only=False).astype(dtype).view(np.uint32))


@pytest.mark.parametrize("dtype,thr", [(np.uint64, 1 << 62), (np.uint64, 0), (np.uint64, 2**64 - 1),
                                       (np.int64, -(1 << 40)), (np.int64, 0), (np.int64, 2**63 - 1),
                                       (np.float64, 0.5), (np.float64, -0.0), (np.float64, np.inf)])
def test_typed_filter_64_matches_arrow(dtype, thr):
    """The restatement the 64-bit GPU filter is compared with (tests/test_gpu_filter64.py): numpy's `<` in
    the column's type, nulls dropped — against Arrow's filter(less(column, scalar)), bit pattern by bit
    pattern (NaN never passes, -0.0 stays -0.0, -0.0 < 0.0 is false)."""
    rng = np.random.default_rng(6)
    n = 20_000
    if dtype == np.float64:
        v = rng.standard_normal(n) * 3
        v[::97] = np.nan
        v[::101] = np.inf
        v[::103] = -np.inf
        v[::107] = -0.0
        v[::109] = 0.0
    else:
        info = np.iinfo(dtype)
        v = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    valid = rng.random(n) > 0.2
    arr = pa.array(v, mask=~valid)
    exp = pc.filter(arr, pc.less(arr, pa.scalar(thr, arr.type)))
    got = oracle.filter_lt_typed(v, valid, thr)
    assert got.dtype == v.dtype and np.array_equal(got.view(np.uint64),
                                                   exp.to_numpy(zero_copy_only=False).astype(dtype).view(np.uint64))


@pytest.mark.parametrize("n,null_frac", [(0, 0.0), (3, 1.0), (1000, 0.2), (70_001, 0.5)])
def test_int32_aggregates_match_arrow(n, null_frac):
    rng = np.random.default_rng(300 + n)
    v = rng.integers(-2**31, 2**31 - 1, size=n, dtype=np.int32, endpoint=True)
    valid = rng.random(n) >= null_frac
    arr = pa.array(v, type=pa.int32(), mask=~valid)
    got = oracle.aggr_nullable(v, valid)
    mm = pc.min_max(arr)
    assert got == {"sum": pc.sum(arr).as_py(), "count": pc.count(arr).as_py(),   # Arrow sums int32 into int64
                   "min": mm["min"].as_py(), "max": mm["max"].as_py()}


@pytest.mark.parametrize("dtype", [np.uint64, np.int64])
@pytest.mark.parametrize("n,null_frac", [(0, 0.0), (1, 0.0), (7, 1.0), (1000, 0.3), (50_000, 0.0)])
def test_64bit_aggregates_match_arrow(dtype, n, null_frac):
    """uint64 / int64 columns: Arrow's sum keeps the column's type and wraps mod 2^64."""
    rng = np.random.default_rng(500 + n)
    info = np.iinfo(dtype)
    v = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)  # full range: the sum wraps
    valid = rng.random(n) >= null_frac
    arr = pa.array(v, type=pa.from_numpy_dtype(dtype), mask=~valid)
    got = oracle.aggr_nullable(v, valid)
    mm = pc.min_max(arr)
    assert got == {"sum": pc.sum(arr).as_py(), "count": pc.count(arr).as_py(),
                   "min": mm["min"].as_py(), "max": mm["max"].as_py()}
