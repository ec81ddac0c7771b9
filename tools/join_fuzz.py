#!/usr/bin/env python
"""join_fuzz.py — long randomised parity run of the join against the CPU oracle (not part of the
pytest suite: minutes, not seconds). Rare events are the point: rows whose two candidate buckets
both overflowed, chains that run through the first candidate, build chunks, hash-space slices.

    python tools/join_fuzz.py [--seeds 40] [--max-rows 3000000]
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
    p.add_argument("--seeds", type=int, default=40)
    p.add_argument("--max-rows", type=int, default=3_000_000)
    a = p.parse_args()
    ctx = Context(0)
    dev = lambda x: (torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda() if len(x)
                     else torch.empty(0, dtype=torch.int32, device="cuda"))
    bad = 0
    t0 = time.time()
    for seed in range(a.seeds):
        rng = np.random.default_rng(90_000 + seed)
        nr = int(rng.integers(1, a.max_rows))
        nl = int(rng.integers(1, a.max_rows))
        kind = seed % 5
        if kind == 0:      # unique build keys, all probes hit
            pk = rng.permutation(nr).astype(np.uint32)
            fk = rng.integers(0, nr, size=nl, dtype=np.uint32)
        elif kind == 1:    # unique sparse keys over the whole 32-bit range, half the probes miss
            pk = rng.choice(np.arange(0, 2**32, max(1, 2**32 // (2 * nr)), dtype=np.uint64)[:2 * nr], nr, replace=False).astype(np.uint32)
            fk = np.where(rng.random(nl) < 0.5, pk[rng.integers(0, nr, size=nl)], rng.integers(0, 2**32, size=nl, dtype=np.uint32)).astype(np.uint32)
        elif kind == 2:    # 2-8 duplicates per build key
            dom = max(1, nr // int(rng.integers(2, 9)))
            pk = rng.integers(0, dom, size=nr, dtype=np.uint32)
            nl = min(nl, 600_000)
            fk = rng.integers(0, dom + dom // 4 + 1, size=nl, dtype=np.uint32)
        elif kind == 3:    # a few very hot keys on the build side (chunked partitions, long chains)
            nr = min(nr, 300_000)
            hot = rng.integers(0, 2**32, size=4, dtype=np.uint32)
            pk = np.where(rng.random(nr) < 0.2, hot[rng.integers(0, 4, size=nr)], rng.integers(0, 2**32, size=nr, dtype=np.uint32)).astype(np.uint32)
            nl = min(nl, 2000)
            fk = np.where(rng.random(nl) < 0.01, hot[rng.integers(0, 4, size=nl)], pk[rng.integers(0, nr, size=nl)]).astype(np.uint32)
        else:              # sequential keys (the benchmark's shape) with a sliced workspace
            pk = np.arange(nr, dtype=np.uint32)
            fk = rng.integers(0, nr, size=nl, dtype=np.uint32)
        x = rng.integers(0, 2**32, size=pk.size, dtype=np.uint32)
        y = rng.integers(0, 2**32, size=fk.size, dtype=np.uint32)
        skip = int(rng.choice([0, 0, 0, 1, 3]))
        if skip:
            kl = oracle.partition_ids(fk, 1 << skip) == 0
            kr = oracle.partition_ids(pk, 1 << skip) == 0
            fk, y, pk, x = fk[kl], y[kl], pk[kr], x[kr]
        exp = oracle.sort_rows(*oracle.join(fk, y, pk, x))
        ws = None
        if kind == 4 and skip == 0 and fk.size and pk.size:
            full, small = ctx.join_ws_bytes(fk.size, pk.size), ctx.join_min_ws_bytes(fk.size, pk.size)
            ws = torch.empty((full + 2 * small) // 3 + 256, dtype=torch.uint8, device="cuda")
        o_fk, o_y, o_x, rows = ctx.join_dev(dev(fk), dev(y), dev(pk), dev(x), out_capacity=max(exp[0].size, 1),
                                            ws=ws, skip_bits=skip)
        agg = ctx.join_aggr_dev(dev(fk), dev(y), dev(pk), dev(x), y_threshold=(1 << 31) if seed % 2 else None,
                                skip_bits=skip)
        torch.cuda.synchronize()
        m = int(rows.cpu().numpy().view(np.uint64)[0])
        h = lambda t: t.cpu().numpy().view(np.uint32)[:m]
        ok = m == exp[0].size
        if ok:
            got = oracle.sort_rows(h(o_fk), h(o_y), h(o_x))
            ok = all(np.array_equal(g, e) for g, e in zip(got, exp))
        ea = oracle.join_aggr(fk, y, pk, x, (1 << 31) if seed % 2 else None)
        ga = agg.cpu().numpy().view(np.uint64)
        ok = ok and {"rows": int(ga[0]), "sum_y": int(ga[1]), "sum_x": int(ga[2])} == ea
        print(f"seed {seed:3d} kind {kind} nl {fk.size:8d} nr {pk.size:8d} out {m:9d} skip {skip} "
              f"{'sliced ' if ws is not None else ''}{'ok' if ok else 'MISMATCH'}", flush=True)
        bad += not ok
    print(f"{a.seeds - bad} of {a.seeds} cases ok in {time.time() - t0:.0f} s")
    ctx.close()
    raise SystemExit(1 if bad else 0)


if __name__ == "__main__":
    main()
