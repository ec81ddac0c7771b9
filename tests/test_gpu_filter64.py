"""Filter over 64-bit columns (b2_filter_lt_64_dev / b2_filter_lt_64_host_into, FilterGpu over uint64 /
int64 / float64 batches; csrc/filter64.cu) against the oracle (oracle.filter_lt_typed) and against Arrow's filter — the reference's
oracle is the Acero plan filter(less(field, literal)) of host/filter/filter_native.cc:52-66. Rows keep
their order, null rows are dropped, NaN rows are never selected."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

_TILE = 2048  # rows per CTA tile of csrc/filter64.cu


@pytest.fixture(params=[0, 1], ids=["single_pass", "two_pass"], autouse=True)
def kernel(request, ctx):
    """Every test runs through both kernels of csrc/filter64.cu (B2_TUNE_FILTER64_KERNEL)."""
    from dpu_olap_b200._lib import TUNE_FILTER64_KERNEL
    ctx.set_tunable(TUNE_FILTER64_KERNEL, request.param)
    yield request.param
    ctx.set_tunable(TUNE_FILTER64_KERNEL, 0)


def _column(rng, dtype, n):
    if dtype == np.float64:
        v = rng.standard_normal(n) * 1e3
        if n:
            k = max(1, n // 17)
            v[rng.integers(0, n, k)] = np.nan
            v[rng.integers(0, n, k)] = -0.0
            v[rng.integers(0, n, k)] = np.inf
            v[rng.integers(0, n, k)] = -np.inf
        return v
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)


def _threshold(rng, dtype, sel):
    if dtype == np.float64:
        return float(np.array([-1e30, -300.0, 0.0, 250.0, 1e30])[int(sel * 4)])
    info = np.iinfo(dtype)
    return int(info.min) + (int(info.max) - int(info.min)) * int(sel * 4) // 4


@pytest.mark.parametrize("dtype", [np.uint64, np.int64, np.float64])
@pytest.mark.parametrize("n", [0, 1, 31, _TILE - 1, _TILE, _TILE + 1, 5 * _TILE, 100_003, (1 << 22) + 77])
@pytest.mark.parametrize("sel", [0.0, 0.25, 1.0])
def test_filter_64_dev_one_batch(ctx, dtype, n, sel):
    rng = np.random.default_rng(n * 3 + int(sel * 8) + np.dtype(dtype).num)
    v = _column(rng, dtype, n)
    thr = _threshold(rng, dtype, sel)
    col = torch.from_numpy(v.view(np.int64).copy()).cuda()
    obuf = torch.full((n + 16,), -7, dtype=torch.int64, device="cuda")
    out, end, total = ctx.filter64_dev(col, dtype, thr, out=obuf[8: 8 + max(n, 1)])
    torch.cuda.synchronize()
    exp = oracle.filter_lt_typed(v, np.ones(n, bool), thr)   # pinned against Arrow in tests/test_oracle_nullable.py
    k = int(total.item())
    assert k == exp.size and int(end[0].item()) == k
    assert np.array_equal(out[:k].cpu().numpy().view(np.uint64), exp.view(np.uint64))   # bit patterns: -0.0 stays -0.0
    # nothing written outside [0, total)
    assert bool((obuf[:8] == -7).all()) and bool((obuf[8 + k:] == -7).all())


@pytest.mark.parametrize("dtype", [np.uint64, np.int64, np.float64])
@pytest.mark.parametrize("lens", [[5], [0, 0, 3, 0], [_TILE, _TILE, _TILE], [1, _TILE - 1, _TILE + 1, 7, 0, 3 * _TILE + 5],
                                  [65536] * 8, [100_000, 1, 33_333, 0, 250_001]])
@pytest.mark.parametrize("nulls", [False, True])
def test_filter_64_dev_ragged_batches_and_nulls(ctx, dtype, lens, nulls):
    rng = np.random.default_rng(sum(lens) + len(lens) + nulls)
    n = sum(lens)
    v = _column(rng, dtype, n)
    thr = _threshold(rng, dtype, 0.5)
    valid = rng.random(n) < 0.7 if nulls else np.ones(n, bool)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = torch.from_numpy(v.view(np.int64).copy()).cuda()
    dvalid = torch.from_numpy(np.packbits(valid, bitorder="little")).cuda() if nulls else None
    out, end, total = ctx.filter64_dev(col, dtype, thr, valid=dvalid, batch_off=torch.from_numpy(off).cuda(),
                                       nbatches=len(lens))
    torch.cuda.synchronize()
    with np.errstate(invalid="ignore"):
        keep = (v < np.array(thr, dtype=dtype)) & valid
    k = int(total.item())
    assert k == int(keep.sum())
    assert np.array_equal(out[:k].cpu().numpy().view(np.uint64), oracle.filter_lt_typed(v, valid, thr).view(np.uint64))
    exp_end = np.concatenate([[0], np.cumsum(keep)])[off[1:]]
    assert np.array_equal(end.cpu().numpy(), exp_end)


def test_filter_64_dev_uniform_batches(ctx):
    rng = np.random.default_rng(11)
    nb, bl = 37, 3000   # boundaries cut tiles
    v = _column(rng, np.int64, nb * bl)
    col = torch.from_numpy(v.copy()).cuda()
    out, end, total = ctx.filter64_dev(col, np.int64, -5, nbatches=nb, batch_len=bl)
    torch.cuda.synchronize()
    keep = v < -5
    assert np.array_equal(out[: int(total.item())].cpu().numpy(), v[keep])
    assert np.array_equal(end.cpu().numpy(), np.cumsum(keep.reshape(nb, bl).sum(axis=1)))


def test_filter_64_dev_rejects_bad_arguments(ctx):
    from dpu_olap_b200.ops import B2Error
    col = torch.zeros(100, dtype=torch.int64, device="cuda")
    with pytest.raises(B2Error):   # batches do not cover the column
        ctx.filter64_dev(col, np.int64, 0, nbatches=3, batch_len=30)
    with pytest.raises(B2Error):   # workspace too small
        ctx.filter64_dev(torch.zeros(1 << 20, dtype=torch.int64, device="cuda"), np.int64, 0,
                         ws=torch.empty(512, dtype=torch.uint8, device="cuda"))
    with pytest.raises(KeyError):  # not a 64-bit type
        ctx.filter64_dev(col, np.uint32, 0)


@pytest.mark.parametrize("dtype,thr", [(np.uint64, 1 << 62), (np.int64, -(1 << 40)), (np.float64, 0.5)])
@pytest.mark.parametrize("nulls", [False, True])
def test_filter_gpu_over_64bit_batches(ctx, dtype, thr, nulls):
    """The operator class over host batches (ragged lengths, Arrow offsets and bitmaps) against
    Arrow's filter, batch by batch."""
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(5 + nulls + np.dtype(dtype).num)
    lens = [65536, 1, 0, 4097, 65536, 12345]
    batches = []
    for n in lens:
        v = _column(rng, dtype, n + 3)
        mask = rng.random(n + 3) < 0.2 if nulls else None
        batches.append(pa.array(v, mask=mask).slice(3, n))   # non-zero Arrow offset
    f = ops.FilterGpu(ctx, batches, threshold=thr)
    f.Prepare()
    got = f.GetResult()
    assert len(got) == len(lens)
    want_rows = 0
    for g, b in zip(got, batches):
        exp = pc.filter(b, pc.less(b, pa.scalar(thr, type=b.type)), null_selection_behavior="drop")
        exp = exp.to_numpy(zero_copy_only=False)
        assert g.dtype == np.dtype(dtype)
        assert np.array_equal(g.view(np.uint64), exp.astype(dtype).view(np.uint64))
        want_rows += len(exp)
    assert f.Run() == want_rows
    assert f.Timers() is not None


def test_filter_64_single_pass_many_tiles_in_flight(ctx, kernel):
    """2^25 rows = 16384 tiles: far more tiles than resident CTAs, so the look-back chain crosses many
    waves; selectivity swings between runs of all-selected and none-selected tiles."""
    n = 1 << 25
    rng = np.random.default_rng(77)
    v = rng.integers(-1000, 1000, size=n, dtype=np.int64)
    v[: n // 4] = -5000                      # a long run of full tiles
    v[n // 2: n // 2 + n // 8] = 5000        # and one of empty tiles
    col = torch.from_numpy(v).cuda()
    out, end, total = ctx.filter64_dev(col, np.int64, 0, nbatches=512, batch_len=n // 512)
    torch.cuda.synchronize()
    keep = v < 0
    k = int(total.item())
    assert k == int(keep.sum())
    assert torch.equal(out[:k], col[torch.from_numpy(keep).cuda()])
    assert np.array_equal(end.cpu().numpy(), np.cumsum(keep.reshape(512, -1).sum(axis=1)))
