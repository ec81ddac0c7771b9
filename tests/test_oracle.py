"""CPU tests: pin the oracle (oracle/olap_oracle.c) before anything is compared against it.

1. generator: against values produced by the REAL libstdc++ <random> + Arrow's vendored PCG
   header (tests/golden/generator_golden.json, made by make_generator_golden.cc);
2. operators: against the reference's own known-answer tests (file:line cited per test);
3. operators at the reference's LargeTest sizes: against Arrow Acero 24 digests
   (tests/golden/arrow_golden.json, made by make_arrow_golden.py).
"""
import hashlib

import numpy as np
import pytest

import oracle


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.uint32).tobytes())
    return h.hexdigest()


# ---- generator ----------------------------------------------------------------------------------
def test_seed_stream(golden):
    g = golden["generator_golden"]
    r = oracle.RandomArrayGenerator(42)
    assert [r.seed() for _ in range(16)] == g["seed_stream_42"]
    r = oracle.RandomArrayGenerator(7)
    assert [r.seed() for _ in range(8)] == g["seed_stream_7"]


def test_generator_arrays(golden):
    for name, v in golden["generator_golden"].items():
        if not isinstance(v, dict):
            continue
        a = oracle.gen_u32(v["seed"] + 1, v["n"], v["lo"], v["hi"])
        assert list(a[:8]) == v["first"], name
        assert int(a[-1]) == v["last"], name
        assert int(a.astype(np.uint64).sum()) == v["sum"], name
        assert int((a < (1 << 30)).sum()) == v["count_lt_2p30"], name
        w = np.arange(1, a.size + 1, dtype=np.uint64)
        assert int(np.bitwise_xor.reduce(a.astype(np.uint64) * w)) == v["xor_weighted"], name


def test_survey_fingerprints():
    # first array of RandomArrayGenerator(42)
    r = oracle.RandomArrayGenerator(42)
    b = r.uint32(65536)
    assert list(b[:3]) == [268, 2955549055, 1465994917]
    assert int((b < (1 << 30)).sum()) == 16358
    assert oracle.sum_u32(b) == 141101534903199
    assert oracle.wang_hash(0) == 0x4636B9C9 and oracle.wang_hash(2) == 0xFF4D1170


# ---- the reference's known-answer tests -------------------------------------------------------------
def test_filter_simple():  # filter_test.cc:24-31
    assert oracle.filter_lt([0, 2, 3, 8, 9]).size == 5


def test_filter_result():  # filter_test.cc:33-61
    keep = {5, 8, 9, 100, 270}
    v = np.array([i if i in keep else i + (1 << 30) for i in range(4096)], dtype=np.uint32)
    assert list(oracle.filter_lt(v)) == [5, 8, 9, 100, 270]


def test_sum_simple():  # aggr_test.cc:24-35
    assert oracle.sum_u32([0, 2, 3, 8, 9]) == 22


def test_take_simple():  # take_test.cc:24-46
    assert list(oracle.take([0, 2, 3, 8, 9], [0, 1, 4])) == [0, 2, 9]


def test_join_simple():  # join_test.cc:40-80
    fk = [0, 2, 3, 8, 9, 10, 12, 13, 18, 19]
    vl = [100, 102, 103, 108, 109, 110, 112, 113, 118, 119]
    pk = [3, 8, 9, 0, 12, 13, 18, 19, 10, 2]
    vr = [53, 58, 59, 50, 62, 63, 68, 69, 60, 52]
    rows = oracle.sort_rows(*oracle.join(fk, vl, pk, vr))
    assert list(rows[0]) == fk
    assert list(rows[1]) == vl
    assert list(rows[2]) == [50, 52, 53, 58, 59, 60, 62, 63, 68, 69]


def test_join_semantics_duplicates_and_misses():
    # Arrow inner-join semantics the DPU path lacks (SURVEY.md §7 hard part 4)
    fk, y = [1, 2, 2, 7], [10, 20, 21, 70]
    pk, x = [2, 2, 1, 5], [200, 201, 100, 500]
    rows = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    assert list(zip(*[r.tolist() for r in rows])) == [
        (1, 10, 100), (2, 20, 200), (2, 20, 201), (2, 21, 200), (2, 21, 201)]


def test_partition_simple():  # partition_test.cc:21-57: {0,2,3,8} -> P=2 sizes 3 / 1
    ids = oracle.partition_ids([0, 2, 3, 8], 2)
    assert sorted(np.bincount(ids, minlength=2).tolist(), reverse=True) == [3, 1]
    assert np.bincount(ids, minlength=2).tolist() == [3, 1]


def test_partition_large_balance():  # partition_test.cc:59-92: 32 partitions within 10 % of the mean
    g = oracle.RandomArrayGenerator(42)
    keys = np.concatenate(oracle.make_random_batches(g, 16, 65536))
    cnt = np.bincount(oracle.partition_ids(keys, 32), minlength=32)
    assert np.all(np.abs(cnt - keys.size / 32) / (keys.size / 32) <= 0.1)


# ---- Arrow Acero digests at the reference's LargeTest sizes ----------------------------------------
def test_filter_vs_arrow(golden):
    ref = golden["arrow_golden"]["filter_128x65536"]
    g = oracle.RandomArrayGenerator(42)
    batches = oracle.make_random_batches(g, 128, 65536)
    outs = [oracle.filter_lt(b) for b in batches]
    assert [o.size for o in outs[:8]] == ref["per_batch_first8"]
    assert sum(o.size for o in outs) == ref["rows"]
    assert sha(*outs) == ref["sha256"]
    assert sum(oracle.sum_u32(b) for b in batches) == golden["arrow_golden"]["sum_128x65536"]["sum"]


def test_take_vs_arrow(golden):
    ref = golden["arrow_golden"]["take_128x65536_8192"]
    g = oracle.RandomArrayGenerator(42)
    vals = oracle.make_random_batches(g, 128, 65536)
    idx = oracle.make_random_batches(g, 128, 8192, 0, 65535)
    outs = [oracle.take(v, i) for v, i in zip(vals, idx)]
    assert sha(*outs) == ref["sha256"]
    assert [int(v) for v in outs[0][:8]] == ref["batch0_first8"]


def test_join_vs_arrow(golden):
    ref = golden["arrow_golden"]["join_128x65536"]
    g = oracle.RandomArrayGenerator(42)
    nb, bs = 128, 65536
    x = np.concatenate(oracle.make_random_batches(g, nb, bs))
    pk = np.concatenate(oracle.make_index_batches(nb, bs))
    y = np.concatenate(oracle.make_random_batches(g, nb, bs))
    fk = np.concatenate(oracle.make_fk_batches(g, bs, nb, bs))
    rows = oracle.sort_rows(*oracle.join(fk, y, pk, x))
    assert rows[0].size == ref["rows"] == nb * bs  # join_test.cc:115-116
    assert sha(*rows) == ref["sorted_sha256"]
    assert oracle.triple_checksum(*rows) == ref["checksum"]


def test_arrow_native_live_small():
    """The Acero plans themselves, live, on a small case (pyarrow is in the image)."""
    pytest.importorskip("pyarrow.acero")
    from oracle import arrow_native as an
    g = oracle.RandomArrayGenerator(42)
    b = oracle.make_random_batches(g, 4, 4096)
    f = an.FilterNative(b, use_threads=False)
    f.Prepare()
    assert np.array_equal(f.GetResult().column(0).to_numpy(), np.concatenate([oracle.filter_lt(x) for x in b]))
    a = an.AggrNative(b)
    a.Prepare()
    assert a.Run() == sum(oracle.sum_u32(x) for x in b)
