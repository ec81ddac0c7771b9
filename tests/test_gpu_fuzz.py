"""Randomised parity sweep on the GPU: many small, odd-shaped cases per operator against the CPU
oracle (seeded, so failures reproduce). Complements the hand-picked edge cases in
test_gpu_dev_ops.py: batch lengths around tile boundaries, ragged layouts with empty batches,
thresholds over the whole range, duplicate / missing / skewed join keys, hash-skip bits."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def _lens(rng, n):
    choices = [0, 1, 3, 4, 5, 4095, 4096, 4097, 8191, 8192, 12288, 65536, 70001]
    return [int(rng.choice(choices)) if rng.random() < 0.7 else int(rng.integers(0, 100_000)) for _ in range(n)]


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_filter(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    thr = int(rng.choice([0, 1, 2**30, 2**31, 2**32 - 1, int(rng.integers(0, 2**32))]))
    if seed % 2 == 0:  # uniform batches
        nb, bl = int(rng.integers(1, 40)), int(rng.choice([1, 5, 4096, 4097, 8192, 16384 + 3, 65536]))
        batches = [rng.integers(0, 2**32, size=bl, dtype=np.uint32) for _ in range(nb)]
        col = dev(np.concatenate(batches))
        out, end, total = ctx.filter_dev(col, nb, bl, thr)
    else:              # ragged batches, some empty
        batches = [rng.integers(0, 2**32, size=n, dtype=np.uint32) for n in _lens(rng, int(rng.integers(1, 25)))]
        if sum(b.size for b in batches) == 0:
            batches.append(rng.integers(0, 2**32, size=7, dtype=np.uint32))
        off = np.concatenate([[0], np.cumsum([b.size for b in batches])]).astype(np.int64)
        col = dev(np.concatenate(batches))
        out, end, total = ctx.filter_ragged_dev(col, off, thr)
    torch.cuda.synchronize()
    exp = [oracle.filter_lt(b, thr) for b in batches]
    n = int(total.cpu()[0])
    assert n == sum(e.size for e in exp)
    assert np.array_equal(host(out)[:n], np.concatenate(exp) if n else np.empty(0, np.uint32))
    assert end.cpu().numpy().tolist() == np.cumsum([e.size for e in exp]).tolist()


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_take_and_sum(ctx, seed):
    rng = np.random.default_rng(2000 + seed)
    nb, vl, il = int(rng.integers(1, 9)), int(rng.integers(1, 300_000)), int(rng.integers(1, 50_000))
    vals = rng.integers(0, 2**32, size=(nb, vl), dtype=np.uint32)
    idx = rng.integers(0, vl, size=(nb, il), dtype=np.uint32)
    got = host(ctx.take_dev(dev(vals.ravel()), vl, dev(idx.ravel()), il, nb)).reshape(nb, il)
    for b in range(nb):
        assert np.array_equal(got[b], oracle.take(vals[b], idx[b]))
    s = ctx.sum_dev(dev(vals.ravel()))
    assert int(s.cpu().numpy().view(np.uint64)[0]) == oracle.sum_u32(vals.ravel())


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_join(ctx, seed):
    rng = np.random.default_rng(3000 + seed)
    nr, nl = int(rng.integers(1, 400_000)), int(rng.integers(1, 400_000))
    domain = int(rng.choice([nr, max(nr // 4, 1), nr * 4, 50]))          # duplicates / misses / heavy skew
    pk = rng.integers(0, domain, size=nr, dtype=np.uint32) if seed % 3 else rng.permutation(nr).astype(np.uint32)
    fk = rng.integers(0, max(domain, 1) + domain // 3 + 1, size=nl, dtype=np.uint32)
    if domain == 50:
        pk, fk = pk[:2000], fk[:3000]  # 50 hot keys: keep the quadratic output small
    x = rng.integers(0, 2**32, size=pk.size, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=fk.size, dtype=np.uint32)
    skip = int(rng.choice([0, 0, 1, 3]))
    if skip:  # keep only the rows a rank of a 2^skip-way sharded join would own
        keep_l = oracle.partition_ids(fk, 1 << skip) == 0
        keep_r = oracle.partition_ids(pk, 1 << skip) == 0
        fk, y, pk, x = fk[keep_l], y[keep_l], pk[keep_r], x[keep_r]
    exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    cap = max(exp[0].size, 1)
    e = lambda a: dev(a) if len(a) else torch.empty(0, dtype=torch.int32, device="cuda")
    o_fk, o_y, o_x, rows = ctx.join_dev(e(fk), e(y), e(pk), e(x), out_capacity=cap, skip_bits=skip)
    torch.cuda.synchronize()
    m = int(rows.cpu().numpy().view(np.uint64)[0])
    assert m == exp[0].size
    got = oracle.sort_rows(host(o_fk)[:m], host(o_y)[:m], host(o_x)[:m])
    for a, b in zip(got, exp):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_partition(ctx, seed):
    rng = np.random.default_rng(4000 + seed)
    n, ncols = int(rng.integers(1, 300_000)), int(rng.integers(1, 4))
    nparts = int(rng.choice([1, 2, 32, 1024, 4096]))
    cols = [rng.integers(0, 2**32 if c else n * 2 + 1, size=n, dtype=np.uint32) for c in range(ncols)]
    outs, off = ctx.partition_dev([dev(c) for c in cols], nparts)
    torch.cuda.synchronize()
    off = off.cpu().numpy()
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
    keys = host(outs[0])
    ids = oracle.partition_ids(keys, nparts)
    assert np.array_equal(ids, np.repeat(np.arange(nparts), np.diff(off)))
    # the columns stay aligned row by row: same multiset of rows
    got = np.stack([host(o) for o in outs], axis=1)
    expd = np.stack(cols, axis=1)
    assert np.array_equal(got[np.lexsort(got.T[::-1])], expd[np.lexsort(expd.T[::-1])])
