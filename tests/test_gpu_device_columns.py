"""Device-resident columns (b2_col) and the Arrow C Device Data Interface (the
reference's own future-work note, host/dpuext/arrow_utils.h:28-29): operators chained on the device
must give what the host-buffer operators / the oracle give, and an exported ArrowDeviceArray must be a
zero-copy, correctly described view that keeps the memory alive."""
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


def _batches(nb, bl, lo=0, hi=0xFFFFFFFF, seed=42):
    g = oracle.RandomArrayGenerator(seed)
    return oracle.make_random_batches(g, nb, bl, lo, hi)


def test_upload_download_round_trip_and_metadata(ctx, ops):
    batches = [np.arange(5, dtype=np.uint32), np.zeros(0, np.uint32), np.arange(7, 20, dtype=np.uint32)]
    col = ops.DeviceColumn.from_host(ctx, batches)
    assert col.rows == 18 and col.nbatches == 3 and col.device_ptr != 0
    assert col.batch_offsets().tolist() == [0, 5, 5, 18]
    for a, b in zip(col.to_host(), batches):
        assert np.array_equal(a, b)
    assert col.sum() == int(sum(int(b.sum()) for b in batches))


@pytest.mark.parametrize("ragged", [False, True])
def test_filter_take_sum_chain_stays_on_the_device(ctx, ops, ragged):
    nb, bl = 24, 65536
    vals = _batches(nb, bl)
    if ragged:
        vals = [v[: bl - 17 * b] for b, v in enumerate(vals)]
    col = ops.DeviceColumn.from_host(ctx, vals)
    f = col.filter_lt(1 << 30)
    exp = [oracle.filter_lt(v) for v in vals]
    assert f.nbatches == nb and f.rows == sum(e.size for e in exp)
    for a, e in zip(f.to_host(), exp):
        assert np.array_equal(a, e)
    # fused check: the sum of the filtered column equals b2_sum_lt's and the oracle's
    assert f.sum() == sum(int(e.sum(dtype=np.uint64)) for e in exp)
    # take the filtered rows by batch-local indices, still on the device
    rng = np.random.default_rng(1)
    idx = [rng.integers(0, max(e.size, 1), size=1000, dtype=np.uint32) for e in exp]
    t = f.take(ops.DeviceColumn.from_host(ctx, idx))
    for a, e, i in zip(t.to_host(), exp, idx):
        assert np.array_equal(a, oracle.take(e, i))


def test_join_on_device_columns(ctx, ops):
    nb, bs = 16, 65536
    g = oracle.RandomArrayGenerator(42)
    x = oracle.make_random_batches(g, nb, bs)
    y = oracle.make_random_batches(g, nb, bs)
    fk = oracle.make_fk_batches(g, bs, nb, bs)
    pk = oracle.make_index_batches(nb, bs)
    cols = [ops.DeviceColumn.from_host(ctx, c) for c in (fk, y, pk, x)]
    o_fk, o_y, o_x = ops.DeviceColumn.join(*cols)
    assert o_fk.rows == nb * bs
    got = oracle.sort_rows(*[np.concatenate(c.to_host()) for c in (o_fk, o_y, o_x)])
    exp = oracle.sort_rows(*oracle.join(*[np.concatenate(c) for c in (fk, y, pk, x)]))
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)
    # chain: the join's x column summed without leaving the device
    assert o_x.sum() == int(exp[2].sum(dtype=np.uint64))
    # duplicate build keys: the output outgrows the first guess and the join re-runs
    pk2 = [p >> 1 << 1 for p in pk]
    cols2 = [ops.DeviceColumn.from_host(ctx, c) for c in (fk, y, pk2, x)]
    outs = ops.DeviceColumn.join(*cols2)
    exp2 = oracle.sort_rows(*oracle.join(*[np.concatenate(c) for c in (fk, y, pk2, x)]))
    got2 = oracle.sort_rows(*[np.concatenate(c.to_host()) for c in outs])
    assert got2[0].size == exp2[0].size
    for a, b in zip(got2, exp2):
        assert np.array_equal(a, b)


def test_arrow_device_array_export_is_a_zero_copy_view(ctx, ops):
    vals = _batches(4, 4096)
    col = ops.DeviceColumn.from_host(ctx, vals)
    a = col.export_arrow()
    assert a.device_type == ops.ARROW_DEVICE_CUDA and a.device_id == ctx.device
    assert a.array.length == col.rows and a.array.null_count == 0 and a.array.offset == 0
    assert a.array.n_buffers == 2 and a.array.n_children == 0
    assert a.array.buffers[0] is None and a.array.buffers[1] == col.device_ptr   # no bitmap; the data IS the column
    assert a.sync_event  # cudaEvent_t*
    # a consumer on the same GPU reads the buffer through the pointer (torch stands in for it)
    ptr, n = a.array.buffers[1], a.array.length
    host = np.empty(n, dtype=np.uint32)
    torch.cuda.synchronize()
    assert ops._lib.lib()  # keep the library referenced
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")
    assert cudart.cudaMemcpy(ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(4 * n), 2) == 0
    assert np.array_equal(host, np.concatenate(vals))
    # the exported array keeps the memory alive after the handle is gone ...
    col.close()
    assert cudart.cudaMemcpy(ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(4 * n), 2) == 0
    assert np.array_equal(host, np.concatenate(vals))
    # ... and importing it back moves ownership: release becomes NULL, the column sees the same bytes
    back = ops.DeviceColumn.import_arrow(ctx, a, [4096] * 4)
    assert not a.array.release
    assert back.device_ptr == ptr and back.sum() == int(np.concatenate(vals).sum(dtype=np.uint64))
    f = back.filter_lt(1 << 31)
    for got, v in zip(f.to_host(), vals):
        assert np.array_equal(got, oracle.filter_lt(v, 1 << 31))
    back.close()
