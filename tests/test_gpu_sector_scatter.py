"""The whole-sector scatter kernel (csrc/partition.cu, part_scatter_sectors_kernel) is only selected
for large fan-outs, which the small parity cases never reach. These cases force it for EVERY
fan-out and push the same partition / join parity checks through it."""
import numpy as np
import pytest
import torch

import oracle
from dpu_olap_b200._lib import TUNE_SCATTER_SECTOR_TILE, TUNE_SCATTER_SECTORS_MIN_BITS
from test_gpu_dev_ops import check_join, check_partition, dev, host, run_join

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 2, 3], ids=["tile8192", "tile16384", "quads", "bulk"])
def sectors(ctx, request):
    default = ctx.get_tunable(TUNE_SCATTER_SECTORS_MIN_BITS)
    tile = ctx.get_tunable(TUNE_SCATTER_SECTOR_TILE)
    ctx.set_tunable(TUNE_SCATTER_SECTORS_MIN_BITS, 0)  # from a fan-out of 2^0: always
    ctx.set_tunable(TUNE_SCATTER_SECTOR_TILE, request.param)  # both shapes of the kernel
    yield ctx
    ctx.set_tunable(TUNE_SCATTER_SECTORS_MIN_BITS, default)
    ctx.set_tunable(TUNE_SCATTER_SECTOR_TILE, tile)


@pytest.mark.parametrize("n,nparts,ncols", [(0, 4, 2), (1, 1, 1), (3, 2, 2), (1000, 2, 3), (8192, 1024, 2),
                                            (8193, 32, 2), (8191, 512, 2), (300_000, 1024, 2),
                                            (1_000_003, 64, 3), (1_000_003, 512, 2), (2_000_000, 4096, 2),
                                            (3_000_000, 1 << 14, 2), (4_500_001, 1 << 19, 2)])
def test_partition_shapes_sector_kernel(sectors, n, nparts, ncols):
    rng = np.random.default_rng(n + nparts)
    cols = [rng.integers(0, 2**32, size=n, dtype=np.uint32) for _ in range(ncols)]
    check_partition(sectors, cols, nparts)


def test_partition_skew_and_skip_bits_sector_kernel(sectors):
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 50, size=500_000, dtype=np.uint32)  # 50 distinct keys: runs far longer than a tile
    check_partition(sectors, [keys, np.arange(keys.size, dtype=np.uint32)], 256)
    keys = np.zeros(100_003, np.uint32)                        # one bucket takes everything
    check_partition(sectors, [keys, np.arange(keys.size, dtype=np.uint32)], 1024)
    keys = rng.integers(0, 2**32, size=200_000, dtype=np.uint32)
    check_partition(sectors, [keys, keys ^ 0x5A5A5A5A], 128, skip_bits=3)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_partition_sector_kernel(sectors, seed):
    rng = np.random.default_rng(7000 + seed)
    n, ncols = int(rng.integers(1, 600_000)), int(rng.integers(1, 4))
    nparts = int(rng.choice([1, 2, 32, 512, 1024, 4096, 1 << 16]))
    cols = [rng.integers(0, 2**32 if c else n * 2 + 1, size=n, dtype=np.uint32) for c in range(ncols)]
    check_partition(sectors, cols, nparts)


@pytest.mark.parametrize("nl,nr", [(1, 1), (1000, 10), (100_000, 100_000), (1_500_000, 700_000)])
def test_join_sector_kernel(sectors, nl, nr):
    rng = np.random.default_rng(nl * 31 + nr)
    pk = rng.permutation(nr).astype(np.uint32)
    fk = rng.integers(0, nr + nr // 5 + 1, size=nl, dtype=np.uint32)  # some probe rows miss
    check_join(sectors, fk, rng.integers(0, 2**32, size=nl, dtype=np.uint32), pk,
               rng.integers(0, 2**32, size=nr, dtype=np.uint32))


def test_join_sliced_small_workspace_sector_kernel(sectors):
    # a workspace below the one-slice need forces hash-space slices: the selection predicate and the
    # slice-capacity checks of the scatter kernel
    rng = np.random.default_rng(11)
    n = 600_000
    pk = rng.permutation(n).astype(np.uint32)
    fk = rng.integers(0, n, size=n, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    small = int(sectors.join_min_ws_bytes(n, n))
    assert small < sectors.join_ws_bytes(n, n)
    check_join(sectors, fk, y, pk, x, ws_bytes=small)


def test_sector_kernel_matches_default_kernel_bytewise(sectors):
    # same input through both kernels: the partitioned (key, value) stream is identical up to the
    # order inside a partition, and the boundaries are identical
    rng = np.random.default_rng(5)
    n = 2_000_003
    cols = [rng.integers(0, 2**32, size=n, dtype=np.uint32), np.arange(n, dtype=np.uint32)]
    a, off_a = sectors.partition_dev([dev(c) for c in cols], 1024)
    sectors.set_tunable(TUNE_SCATTER_SECTORS_MIN_BITS, 99)
    b, off_b = sectors.partition_dev([dev(c) for c in cols], 1024)
    torch.cuda.synchronize()
    assert np.array_equal(off_a.cpu().numpy(), off_b.cpu().numpy())
    off = off_a.cpu().numpy()
    ka, va, kb, vb = host(a[0]), host(a[1]), host(b[0]), host(b[1])
    part = np.repeat(np.arange(1024), np.diff(off))
    oa, ob = np.lexsort((va, part)), np.lexsort((vb, part))
    assert np.array_equal(ka[oa], kb[ob]) and np.array_equal(va[oa], vb[ob])


@pytest.mark.parametrize("n,parts", [(1, 1024), (5, 512), (8191, 1024), (100_003, 1024), (1_000_001, 512), (70_001, 4)])
def test_sector_kernel_stays_inside_its_buffers(sectors, n, parts):
    """Output pairs, boundaries and workspace sit between sentinel guards (the pool has no
    compute-sanitizer): nothing outside [0, n) rows is written, every row is written exactly once."""
    from test_gpu_dev_ops import GUARD, _guarded, _guards_intact
    S64 = -0x0123456789ABCDEF
    rng = np.random.default_rng(n + parts)
    key = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    val = np.arange(n, dtype=np.uint32)
    g_p, pairs = _guarded(n, torch.int64, S64)
    g_o, off = _guarded(parts + 1, torch.int64, S64)
    ws_n = int(sectors._lib.b2_shuffle_ws_bytes(n, parts)) + 512
    g_w, ws = _guarded(ws_n, torch.uint8, 0x5A)
    assert pairs.data_ptr() % 32 == 0  # else the plain kernel would be chosen
    sectors.shuffle_partition_dev(dev(key), dev(val), parts, pairs_out=pairs, dest_off=off, ws=ws)
    torch.cuda.synchronize()
    assert _guards_intact(g_p, n, S64) and _guards_intact(g_o, parts + 1, S64) and _guards_intact(g_w, ws_n, 0x5A)
    got = pairs.cpu().numpy().view(np.uint64)
    gk, gv = (got & 0xFFFFFFFF).astype(np.uint32), (got >> 32).astype(np.uint32)
    assert np.array_equal(np.sort(gv), val)            # a permutation: every row exactly once
    assert np.array_equal(gk, key[gv])                 # the key travelled with its value
    o = off.cpu().numpy()
    assert np.array_equal(oracle.partition_ids(gk, parts), np.repeat(np.arange(parts, dtype=np.uint32), np.diff(o)))
