"""Take over 64-bit values with 32-bit indices (b2_take_64_dev / b2_take_64_host, TakeGpu over
uint64 / int64 batches) against numpy's / Arrow's take, batch-local as take_native.cc:24-31."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nb,vl,il,misalign", [(1, 1, 1, 0), (3, 1000, 37, 0), (8, 65536, 8192, 0), (4, 4096, 10_000, 0),
                                                (5, 777, 1024, 1), (2, 1 << 20, 1 << 18, 0), (7, 100, 0, 0)])
def test_take_64_dev(ctx, nb, vl, il, misalign):
    rng = np.random.default_rng(nb * 7 + vl + il)
    vals = rng.integers(0, 2**64 - 1, size=(nb, vl), dtype=np.uint64, endpoint=True)
    idx = rng.integers(0, vl, size=(nb, il), dtype=np.uint32)
    dv = torch.from_numpy(vals.view(np.int64).ravel()).cuda()
    # misalign: indices and output start 4 / 8 bytes off a 16-byte boundary (row-by-row path)
    ibuf = torch.from_numpy(np.concatenate([np.zeros(misalign, np.int32), idx.view(np.int32).ravel()])).cuda()
    obuf = torch.full((nb * il + misalign + 8,), -1, dtype=torch.int64, device="cuda")
    out = obuf[misalign: misalign + nb * il]
    ctx.take64_dev(dv, vl, ibuf[misalign:], il, nb, out=out)
    torch.cuda.synchronize()
    exp = np.take_along_axis(vals, idx.astype(np.int64), axis=1)
    assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(nb, il), exp)
    assert bool((obuf[:misalign] == -1).all()) and bool((obuf[misalign + nb * il:] == -1).all())  # guards


@pytest.mark.parametrize("dtype", [np.uint64, np.int64])
@pytest.mark.parametrize("shapes", [[(1, 1)], [(1000, 37), (1000, 37), (5, 12)], [(65536, 8192)] * 4,
                                    [(10, 0), (3, 7), (4096, 4096), (4096, 4096)]])
def test_take_gpu_over_64bit_batches(ctx, dtype, shapes):
    """The operator class (host buffers) against Arrow's take per batch; ragged batch shapes."""
    from dpu_olap_b200 import ops
    rng = np.random.default_rng(len(shapes) + shapes[0][0] + (dtype == np.int64))
    info = np.iinfo(dtype)
    vals = [rng.integers(info.min, info.max, size=vl, dtype=dtype, endpoint=True) for vl, _ in shapes]
    idx = [rng.integers(0, vl, size=il, dtype=np.uint32) for vl, il in shapes]
    t = ops.TakeGpu(ctx, [pa.array(v) for v in vals], [pa.array(i) for i in idx])
    t.Prepare()
    got = t.Run()
    assert len(got) == len(shapes)
    for g, v, i in zip(got, vals, idx):
        exp = pc.take(pa.array(v), pa.array(i), boundscheck=False).to_numpy()
        assert g.dtype == dtype and np.array_equal(g, exp)


def test_take_64_rejects_nullable_and_mixed(ctx):
    from dpu_olap_b200 import ops
    v = pa.array([1, None, 3], type=pa.uint64())
    with pytest.raises(TypeError):
        ops.TakeGpu(ctx, [v], [pa.array(np.array([0, 2], np.uint32))])
    with pytest.raises(TypeError):
        ops.TakeGpu(ctx, [np.zeros(4, np.uint32), np.zeros(4, np.uint64)], [np.zeros(1, np.uint32)] * 2)
