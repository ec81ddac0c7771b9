"""bench.py's full-size self-checks are torch restatements of oracle functions: pin them to the
oracle on the CPU so that a green bench line means what it says (the SF=2048 runs
compared counts, not content)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import oracle  # noqa: E402


def _as_i32(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32))


@pytest.mark.parametrize("n", [0, 1, 1000, 70001])
def test_triple_checksum_matches_oracle(n):
    rng = np.random.default_rng(n)
    a, b, c = (rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32) for _ in range(3))
    if n > 10:  # the ends of the range
        a[0], b[1], c[2] = 0xFFFFFFFF, 0x80000000, 0
    want = oracle.triple_checksum(a, b, c)
    assert bench.triple_checksum_torch(_as_i32(a), _as_i32(b), _as_i32(c), chunk=4096) == want
    # order independent, content dependent
    p = rng.permutation(n)
    assert bench.triple_checksum_torch(_as_i32(a[p]), _as_i32(b[p]), _as_i32(c[p])) == want
    if n > 10:
        c2 = c.copy()
        c2[[3, 4]] = c2[[4, 3]]  # same column multiset, wrong pairing
        if c2[3] != c2[4]:
            assert bench.triple_checksum_torch(_as_i32(a), _as_i32(b), _as_i32(c2)) != want


def test_filter_content_check_accepts_oracle_result_and_rejects_damage():
    nb, B = 5, bench.FILTER_BATCH
    g = oracle.RandomArrayGenerator(42)
    col = np.concatenate(oracle.make_random_batches(g, nb, B))
    thr = 1 << 30
    out = oracle.filter_lt(col, thr)
    ends = np.cumsum([(col[b * B:(b + 1) * B] < thr).sum() for b in range(nb)]).astype(np.int64)
    t_col, t_out, t_end = _as_i32(col), _as_i32(out), torch.from_numpy(ends)
    assert bench.check_filter_content(t_col, nb, t_out, t_end, thr, "t") == out.size
    bad = t_out.clone()
    bad[out.size // 2] ^= 1
    with pytest.raises(SystemExit):
        bench.check_filter_content(t_col, nb, bad, t_end, thr, "t")
    bad_end = t_end.clone()
    bad_end[2] += 1
    with pytest.raises(SystemExit):
        bench.check_filter_content(t_col, nb, t_out, bad_end, thr, "t")


def test_ncu_traffic_tool_reads_a_committed_capture(tmp_path):
    import json
    import subprocess
    csv = ROOT / "profiles" / "r1_launches_join_sf1024_after_sector_scatter.csv"
    out = tmp_path / "t.json"
    subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_traffic.py"), "--join", str(csv), "--join-rows", "2**31",
                    "--out", str(out)], check=True, capture_output=True)
    j = json.loads(out.read_text())["join"]
    assert 100 < j["dram_bytes_per_row"] < 130 and j["launches"] == 19
