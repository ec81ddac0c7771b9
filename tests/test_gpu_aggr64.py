"""uint64 / int64 aggregates (b2_aggr_64_dev / b2_aggr_64_host, SumGpu over 64-bit columns) against
the oracle and against Arrow's compute kernels: sum in the column's type (wrapping mod 2^64), count,
min / max in the type's order, null rows skipped."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest
import torch

import oracle
from test_gpu_nullable import dev_bits

pytestmark = pytest.mark.gpu


def _values(rng, n, dtype, full_range=True):
    info = np.iinfo(dtype)
    if full_range:
        return rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    return rng.integers(0, 1000, size=n, dtype=dtype)


@pytest.mark.parametrize("dtype", [np.uint64, np.int64])
@pytest.mark.parametrize("n,null_frac,misalign", [(0, 0.0, 0), (1, 0.0, 0), (2, 0.0, 1), (5, 1.0, 0), (1000, 0.3, 0),
                                                    (1 << 20, 0.5, 0), ((1 << 21) + 13, 0.05, 0),
                                                    (100_001, 0.5, 1), (100_000, 0.0, 1), (70_000, 0.999, 0)])
def test_aggregates_64_dev(ctx, dtype, n, null_frac, misalign):
    from dpu_olap_b200.ops import decode_aggr
    rng = np.random.default_rng(n + misalign + (dtype == np.int64))
    v = _values(rng, n, dtype)
    valid = rng.random(n) >= null_frac
    buf = torch.from_numpy(np.concatenate([np.zeros(misalign, dtype), v]).view(np.int64)).cuda()
    col = buf[misalign:]  # values 8 bytes off a 16-byte boundary when misalign == 1
    got = decode_aggr(ctx.aggr_dev(col, dev_bits(valid), dtype=dtype), dtype)
    assert got == oracle.aggr_nullable(v, valid)
    got = decode_aggr(ctx.aggr_dev(col, None, dtype=dtype), dtype)  # no bitmap: every row counts
    assert got == oracle.aggr_nullable(v, np.ones(n, bool))


@pytest.mark.parametrize("dtype", [np.uint64, np.int64])
def test_aggregates_64_extremes(ctx, dtype):
    """Wrap-around sums and the ends of the range as min / max."""
    from dpu_olap_b200.ops import decode_aggr
    info = np.iinfo(dtype)
    v = np.array([info.max, info.max, info.min, 1, info.max], dtype=dtype)
    t = torch.from_numpy(v.view(np.int64)).cuda()
    got = decode_aggr(ctx.aggr_dev(t, None, dtype=dtype), dtype)
    arr = pa.array(v)
    mm = pc.min_max(arr)
    assert got == {"sum": pc.sum(arr).as_py(), "count": 5, "min": mm["min"].as_py(), "max": mm["max"].as_py()}
    assert got["min"] == info.min and got["max"] == info.max


@pytest.mark.parametrize("dtype", [np.uint64, np.int64])
@pytest.mark.parametrize("lens,null_frac", [([1], 0.0), ([5, 0, 3], 1.0), ([1000, 37, 65536], 0.3),
                                            ([1 << 18] * 4, 0.0), ([100_001, 1, 70_000], 0.5)])
def test_sum_gpu_over_64bit_arrow_batches(ctx, dtype, lens, null_frac):
    """The operator class (host buffers, upload inside the call) against Arrow itself, batches with
    their own validity bitmaps and a sliced (offset) batch."""
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(sum(lens) + (dtype == np.int64))
    typ = pa.from_numpy_dtype(dtype)
    batches = []
    for i, n in enumerate(lens):
        v = _values(rng, n + 3, dtype)
        valid = rng.random(n + 3) >= null_frac
        arr = pa.array(v, type=typ, mask=~valid) if null_frac else pa.array(v, type=typ)
        batches.append(arr.slice(3 if i % 2 else 0, n))  # odd batches: offset 3 into their buffers
    whole = pa.chunked_array(batches, type=typ)
    s = ops.SumGpu(ctx, batches)
    s.Prepare()
    mm = pc.min_max(whole)
    assert s.Aggregates() == {"sum": pc.sum(whole).as_py(), "count": pc.count(whole).as_py(),
                              "min": mm["min"].as_py(), "max": mm["max"].as_py()}
    assert s.Run() == pc.sum(whole).as_py()


def test_sum_gpu_rejects_mixed_widths(ctx):
    from dpu_olap_b200 import ops
    with pytest.raises(TypeError):
        ops.SumGpu(ctx, [np.zeros(4, np.uint32), np.zeros(4, np.uint64)])
    with pytest.raises(TypeError):
        ops.FilterGpu(ctx, [np.zeros(4, np.uint32), np.zeros(4, np.uint64)])   # one width per operator
    assert ops.FilterGpu(ctx, [np.zeros(4, np.uint64)]).dtype == np.dtype(np.uint64)   # 64-bit filters: csrc/filter64.cu
