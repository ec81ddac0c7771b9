"""The perfect-hash path of the probe kernel (csrc/join.cu: entry = low hash bits, one 64-bit exchange per
build row, one 64-bit load per probe row) is only planned for >= 2^29 build rows by default. These
cases force it at sizes the oracle finishes in seconds (B2_TUNE_JOIN_DIRECT_MIN_ROWS = 1: the planner
takes 19 partition bits whenever the hash has them) and push the join parity checks through it: unique
keys, missing keys, duplicate build keys (the partition falls back to the bucketised table), hash-space
slices, skip bits, the fused aggregates."""
import numpy as np
import pytest
import torch

import oracle
from dpu_olap_b200._lib import TUNE_JOIN_DIRECT_MIN_ROWS
from test_gpu_dev_ops import check_join, dev, host

pytestmark = pytest.mark.gpu

EXTRA = 96 << 20  # the forced plan keeps 2^19 partition boundaries per side


@pytest.fixture
def direct(ctx):
    old = ctx.get_tunable(TUNE_JOIN_DIRECT_MIN_ROWS)
    ctx.set_tunable(TUNE_JOIN_DIRECT_MIN_ROWS, 1)
    yield ctx
    ctx.set_tunable(TUNE_JOIN_DIRECT_MIN_ROWS, old)


def ws_for(ctx, nl, nr):
    return int(ctx.join_ws_bytes(nl, nr)) + EXTRA


def test_reference_simple_test_direct(direct):  # join_test.cc:40-80
    fk = np.array([0, 2, 3, 8, 9, 10, 12, 13, 18, 19], np.uint32)
    pk = np.array([3, 8, 9, 0, 12, 13, 18, 19, 10, 2], np.uint32)
    check_join(direct, fk, fk + 100, pk, pk + 50, ws_bytes=ws_for(direct, 10, 10))


@pytest.mark.parametrize("nl,nr", [(1, 1), (1000, 10), (100_000, 100_000), (1_500_000, 700_000), (300_000, 2_000_000)])
def test_unique_keys_with_misses_direct(direct, nl, nr):
    rng = np.random.default_rng(nl * 31 + nr)
    pk = rng.permutation(nr).astype(np.uint32)
    fk = rng.integers(0, nr + nr // 5 + 1, size=nl, dtype=np.uint32)  # some probe rows miss
    check_join(direct, fk, rng.integers(0, 2**32, size=nl, dtype=np.uint32), pk,
               rng.integers(0, 2**32, size=nr, dtype=np.uint32), ws_bytes=ws_for(direct, nl, nr))


def test_full_range_keys_direct(direct):
    rng = np.random.default_rng(3)
    pk = np.unique(rng.integers(0, 2**32, size=400_000, dtype=np.uint32))
    pk = np.concatenate([pk, np.array([0, 1, 2**32 - 1], np.uint32)])  # the empty-entry markers are legal keys
    pk = np.unique(pk)
    rng.shuffle(pk)
    fk = np.concatenate([rng.choice(pk, 300_000), rng.integers(0, 2**32, size=100_000, dtype=np.uint32),
                         np.array([0, 1, 0, 1, 2**32 - 1], np.uint32)])
    check_join(direct, fk, rng.integers(0, 2**32, size=fk.size, dtype=np.uint32), pk,
               rng.integers(0, 2**32, size=pk.size, dtype=np.uint32), ws_bytes=ws_for(direct, fk.size, pk.size))


def test_duplicate_build_keys_fall_back_direct(direct):
    rng = np.random.default_rng(4)
    pk = rng.integers(0, 50_000, size=200_000, dtype=np.uint32)  # every key ~4 times
    fk = rng.integers(0, 60_000, size=150_000, dtype=np.uint32)
    check_join(direct, fk, rng.integers(0, 2**32, size=fk.size, dtype=np.uint32), pk,
               np.arange(pk.size, dtype=np.uint32), ws_bytes=ws_for(direct, fk.size, pk.size))
    # a few duplicates among mostly unique keys: single partitions fall back, the rest stays on the fast path
    pk = np.concatenate([rng.permutation(300_000).astype(np.uint32), np.array([7, 7, 7, 123_456], np.uint32)])
    fk = rng.integers(0, 300_000, size=400_000, dtype=np.uint32)
    fk[:1000] = 7
    check_join(direct, fk, rng.integers(0, 2**32, size=fk.size, dtype=np.uint32), pk,
               np.arange(pk.size, dtype=np.uint32), ws_bytes=ws_for(direct, fk.size, pk.size))


def test_hot_key_and_skip_bits_direct(direct):
    rng = np.random.default_rng(5)
    pk = rng.permutation(500_000).astype(np.uint32)
    fk = np.where(rng.random(600_000) < 0.5, 42, rng.integers(0, 500_000, size=600_000)).astype(np.uint32)
    y = rng.integers(0, 2**32, size=fk.size, dtype=np.uint32)
    x = rng.integers(0, 2**32, size=pk.size, dtype=np.uint32)
    check_join(direct, fk, y, pk, x, ws_bytes=ws_for(direct, fk.size, pk.size))
    check_join(direct, fk, y, pk, x, ws_bytes=ws_for(direct, fk.size, pk.size), skip_bits=3)


def test_sliced_small_workspace_direct(direct):
    rng = np.random.default_rng(11)
    n = 600_000
    pk = rng.permutation(n).astype(np.uint32)
    fk = rng.integers(0, n, size=n, dtype=np.uint32)
    from dpu_olap_b200._lib import B2Error
    y = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    full, passed, refused = ws_for(direct, n, n), 0, 0
    for frac in (0.6, 0.45, 0.3, 0.2, 0.12):  # ever smaller workspaces: more and more slices, then a refusal
        try:
            check_join(direct, fk, y, pk, x, ws_bytes=int(full * frac))
            passed += 1
        except B2Error as e:
            assert "workspace" in str(e).lower()
            refused += 1
    assert passed >= 2


def test_join_aggregate_direct(direct):
    rng = np.random.default_rng(6)
    nr, nl = 400_000, 500_000
    pk = rng.permutation(nr).astype(np.uint32)
    x = rng.integers(0, 2**32, size=nr, dtype=np.uint32)
    fk = rng.integers(0, nr + 1000, size=nl, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=nl, dtype=np.uint32)
    ws = torch.empty(ws_for(direct, nl, nr), dtype=torch.uint8, device="cuda")
    for thr in (None, 1 << 30):
        res = direct.join_aggr_dev(dev(fk), dev(y), dev(pk), dev(x), y_threshold=thr, ws=ws)
        torch.cuda.synchronize()
        keep = np.ones(nl, bool) if thr is None else y < thr
        e_fk, e_y, e_x = oracle.join(fk[keep], y[keep], pk, x)
        got = res.cpu().numpy().view(np.uint64)
        assert int(got[0]) == e_fk.size
        assert int(got[1]) == int(e_y.astype(np.uint64).sum()) and int(got[2]) == int(e_x.astype(np.uint64).sum())
