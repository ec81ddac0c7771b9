"""b2_set (include/b200olap.h "device set"): the GPUs of one node behind one handle, driven by ONE
process — the counterpart of dpu::DpuSet::allocate(nr_dpus) (dpuext.hpp:704-739). The operator
classes take a DeviceSet in place of a Context; the cases are the reference's GoogleTests again.
A one-member set runs everywhere; the sharded cases need 2 / 4 / 8 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_set.py -m gpu`)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ngpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture(scope="module", params=[1, 2, 4, 8])
def dset(request, built_lib):
    from dpu_olap_b200 import ops
    n = request.param
    if _ngpus() < n:
        pytest.skip(f"needs {n} GPUs")
    s = ops.DeviceSet(n)
    assert len(s) == n and (n == 1 or s.peer_access)
    yield s
    s.close()


def _join_inputs(nb, bs):
    g = oracle.RandomArrayGenerator(42)
    x = oracle.make_random_batches(g, nb, bs)
    y = oracle.make_random_batches(g, nb, bs)
    fk = oracle.make_fk_batches(g, bs, nb, bs)
    pk = oracle.make_index_batches(nb, bs)
    return fk, y, pk, x


def _check_join(dset, fk, y, pk, x):
    from dpu_olap_b200 import ops
    left = [{"fk": a, "y": b} for a, b in zip(fk, y)]
    right = [{"pk": a, "x": b} for a, b in zip(pk, x)]
    j = ops.JoinGpu(dset, left, right)
    j.Prepare()
    out = j.Run()
    cat = lambda c: np.concatenate(c) if len(c) else np.zeros(0, np.uint32)
    exp = oracle.sort_rows(*oracle.join(cat(fk), cat(y), cat(pk), cat(x)))
    got = oracle.sort_rows(out["fk"], out["y"], out["x"])
    assert got[0].size == exp[0].size
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)
    # the fused join -> aggregate pipeline over the same set (b2_set_join_aggr_u32_host): every member adds
    # the payloads of its local join's rows, the host adds the members' sums
    for thr in (None, 1 << 30):
        keep = np.ones(exp[0].size, bool) if thr is None else exp[1] < thr
        want = {"rows": int(keep.sum()), "sum_y": int(exp[1][keep].astype(np.uint64).sum(dtype=np.uint64)),
                "sum_x": int(exp[2][keep].astype(np.uint64).sum(dtype=np.uint64))}
        assert j.RunAggregate(thr) == want
    return out, j


def test_set_filter_sum_take(dset):
    from dpu_olap_b200 import ops
    g = oracle.RandomArrayGenerator(42)
    batches = oracle.make_random_batches(g, 37, 65536)  # 37: does not split evenly over the members
    f = ops.FilterGpu(dset, batches)
    f.Prepare()
    chunks = f.GetResult()
    assert len(chunks) == len(batches)
    for c, b in zip(chunks, batches):
        assert np.array_equal(c, oracle.filter_lt(b))
    assert f.Run() == sum(c.size for c in chunks)
    s = ops.SumGpu(dset, batches)
    s.Prepare()
    assert s.Run() == sum(oracle.sum_u32(b) for b in batches)
    idx = oracle.make_random_batches(g, 37, 8192, 0, 65535)
    t = ops.TakeGpu(dset, batches, idx)
    t.Prepare()
    out = t.Run()
    for b in range(37):
        assert np.array_equal(out[b], oracle.take(batches[b], idx[b]))
    # fewer batches than members, and none at all
    assert ops.SumGpu(dset, batches[:1]).Run() == oracle.sum_u32(batches[0])
    assert ops.SumGpu(dset, []).Run() == 0


def test_set_join_simple_test(dset):  # join_test.cc:40-80
    fk = [np.array([0, 2, 3, 8, 9], np.uint32), np.array([10, 12, 13, 18, 19], np.uint32)]
    y = [np.array([100, 102, 103, 108, 109], np.uint32), np.array([110, 112, 113, 118, 119], np.uint32)]
    pk = [np.array([3, 8, 9, 0, 2], np.uint32), np.array([12, 13, 18, 19, 10], np.uint32)]
    x = [np.array([53, 58, 59, 50, 52], np.uint32), np.array([62, 63, 68, 69, 60], np.uint32)]
    out, _ = _check_join(dset, fk, y, pk, x)
    assert sorted(out["fk"].tolist()) == [0, 2, 3, 8, 9, 10, 12, 13, 18, 19]


def test_set_join_large_test(dset):  # join_test.cc:82-121 at 32 x 65536
    fk, y, pk, x = _join_inputs(32, 65536)
    out, j = _check_join(dset, fk, y, pk, x)
    assert out["fk"].size == 32 * 65536
    assert j.Timers() is not None
    # again on the same set: receive buffers and events are reused
    _check_join(dset, fk, y, pk, x)


def test_set_join_duplicates_misses_and_skew(dset):
    rng = np.random.default_rng(7)
    nb, bs = 8, 20000
    # duplicate build keys (every key 3 times) and probe keys without a match: Arrow inner-join semantics;
    # the output outgrows the first-guess capacity of a member, which must repeat its local join
    pk = [rng.integers(0, 40000, bs, dtype=np.uint32) for _ in range(nb)]
    pk = [np.concatenate([p, p, p])[:bs] for p in pk]
    fk = [rng.integers(0, 60000, bs, dtype=np.uint32) for _ in range(nb)]
    x = [rng.integers(0, 2**32, bs, dtype=np.uint32) for _ in range(nb)]
    y = [rng.integers(0, 2**32, bs, dtype=np.uint32) for _ in range(nb)]
    _check_join(dset, fk, y, pk, x)
    # heavy skew: most probe rows carry ONE key, so one member receives far more than an even share
    # and the set must grow its receive buffers and repeat the exchange
    fk = [np.where(rng.random(bs) < 0.9, np.uint32(77), rng.integers(0, 40000, bs, dtype=np.uint32)).astype(np.uint32)
          for _ in range(nb)]
    pk = [np.arange(b * bs, (b + 1) * bs, dtype=np.uint32) for b in range(nb)]
    _check_join(dset, fk, y, pk, x)
    # empty sides
    _check_join(dset, [], [], pk, x)
    _check_join(dset, fk, y, [], [])


def test_set_rejects_bad_shapes(built_lib):
    import ctypes as C

    from dpu_olap_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.b2_set_create(None, 0, C.byref(h)) == 1            # B2_ERR_INVALID
    two_same = (C.c_int * 2)(0, 0)
    assert lib.b2_set_create(two_same, 2, C.byref(h)) == 1        # a device appears once
    assert lib.b2_set_size(None) == 0 and lib.b2_set_ctx(None, 0) is None
