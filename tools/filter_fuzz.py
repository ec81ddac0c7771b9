#!/usr/bin/env python
"""filter_fuzz.py — randomised parity run of the filter kernel (plain, ragged, nullable, int32 /
float32, carry-in appends) against the CPU oracle; odd batch lengths, tiny and large grids.

    python tools/filter_fuzz.py [--cases 80]
"""
import argparse, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import numpy as np
    import torch
    import oracle
    from dpu_olap_b200.ops import Context
    p = argparse.ArgumentParser()
    p.add_argument("--cases", type=int, default=80)
    a = p.parse_args()
    ctx = Context(0)
    bad = 0
    t0 = time.time()
    for case in range(a.cases):
        rng = np.random.default_rng(70_000 + case)
        nb = int(rng.choice([1, 2, 3, 7, 16, 100, 1000, 5000]))
        bl = int(rng.choice([1, 3, 31, 32, 100, 4095, 4096, 4097, 4100, 8192 + 36, 65536, 65536 + 4, 100_003]))
        while nb * bl > 40_000_000:
            nb = max(1, nb // 2)
        dtype = [np.uint32, np.int32, np.float32][case % 3]
        n = nb * bl
        if dtype == np.float32:
            v = (rng.standard_normal(n) * 2).astype(np.float32)
            v[::53] = np.nan
            thr = float(rng.choice([0.0, -1.0, 1.5, np.inf]))
        elif dtype == np.int32:
            v = rng.integers(-2**31, 2**31 - 1, size=n, dtype=np.int32, endpoint=True)
            thr = int(rng.choice([0, -2**30, 2**30, -2**31, 2**31 - 1]))
        else:
            v = rng.integers(0, 2**32, size=n, dtype=np.uint32)
            thr = int(rng.choice([1 << 30, 0, 1, 0xFFFFFFFF, 42_949_673]))
        null_frac = float(rng.choice([0.0, 0.0, 0.1, 0.9]))
        valid = rng.random(n) >= null_frac
        t = torch.from_numpy(v.view(np.int32)).cuda()
        bits = torch.from_numpy(oracle.pack_bits(valid)).cuda() if null_frac else None
        out, end, total = ctx.filter_typed_dev(t, dtype, thr, nb, bl, valid=bits)
        torch.cuda.synchronize()
        exp = [oracle.filter_lt_typed(v[b * bl:(b + 1) * bl], valid[b * bl:(b + 1) * bl], thr) for b in range(nb)]
        m = int(total.cpu()[0])
        ok = m == sum(e.size for e in exp)
        ok = ok and np.array_equal(end.cpu().numpy()[:nb], np.cumsum([e.size for e in exp]))
        ok = ok and np.array_equal(out.cpu().numpy().view(np.uint32)[:m], np.concatenate(exp).view(np.uint32))
        if dtype == np.uint32 and not null_frac:  # the ragged entry point on the same rows, split unevenly
            cuts = np.sort(rng.integers(0, n + 1, size=min(6, n)))
            off = np.concatenate([[0], cuts, [n]]).astype(np.int64)
            o2, e2, t2 = ctx.filter_ragged_dev(t, off, thr)
            torch.cuda.synchronize()
            m2 = int(t2.cpu()[0])
            ok = ok and m2 == m and np.array_equal(o2.cpu().numpy().view(np.uint32)[:m2], np.concatenate(exp).view(np.uint32))
            per = [int((v[off[i]:off[i + 1]] < np.uint32(thr)).sum()) if thr <= 0xFFFFFFFF else 0 for i in range(off.size - 1)]
            ok = ok and np.array_equal(e2.cpu().numpy()[:off.size - 1], np.cumsum(per))
        print(f"case {case:3d} {np.dtype(dtype).name:8s} nb {nb:5d} bl {bl:7d} nulls {null_frac:.1f} thr {thr!s:>12s} "
              f"selected {m:9d} {'ok' if ok else 'MISMATCH'}", flush=True)
        bad += not ok
    print(f"{a.cases - bad} of {a.cases} cases ok in {time.time() - t0:.0f} s")
    ctx.close()
    raise SystemExit(1 if bad else 0)


if __name__ == "__main__":
    main()
