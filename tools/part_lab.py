#!/usr/bin/env python
"""part_lab.py — time ONE radix pass (histogram + scan + scatter) over n (key, value) rows for a
grid of fan-outs and sizes, per scatter-kernel shape (tuning aid: separates the cost of the
fan-out from the cost of the footprint).

    python tools/part_lab.py [--logn 27,29,31] [--parts 64,256,1024] [--variants 0,8,s]
"""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from dpu_olap_b200.ops import Context
    p = argparse.ArgumentParser()
    p.add_argument("--logn", default="27,29,31")
    p.add_argument("--parts", default="64,256,1024")
    p.add_argument("--variants", default="0,8,s")
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--sector-tile", type=int, default=None, help="B2_TUNE_SCATTER_SECTOR_TILE: 0, 1 or 2 (quads kernel)")
    a = p.parse_args()
    ctx = Context(0)
    if a.sector_tile is not None:
        ctx.set_tunable(4, a.sector_tile)
    for logn in [int(x) for x in a.logn.split(",")]:
        n = 1 << logn
        key = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
        val = torch.arange(n, dtype=torch.int32, device="cuda")
        pairs = torch.empty(n, dtype=torch.int64, device="cuda")
        for parts in [int(x) for x in a.parts.split(",")]:
            off = torch.empty(parts + 1, dtype=torch.int64, device="cuda")
            ws = torch.empty(int(ctx._lib.b2_shuffle_ws_bytes(n, parts)) + 512, dtype=torch.uint8, device="cuda")
            for v in a.variants.split(","):
                # "s" = the whole-sector kernel for every fan-out; a number = that shape of the plain kernel
                ctx.set_tunable(0, 0 if v == "s" else 99)          # B2_TUNE_SCATTER_SECTORS_MIN_BITS
                ctx.set_tunable(2, 0 if v == "s" else int(v))       # B2_TUNE_SCATTER_SHAPE
                step = lambda: ctx.shuffle_partition_dev(key, val, parts, pairs_out=pairs, dest_off=off, ws=ws)
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.reps
                # the values are the row numbers: a permutation, and every row sits in its bucket's range
                ok = int(off[-1]) == n and int((pairs >> 32).sum()) == n * (n - 1) // 2
                print(json.dumps({"log2_rows": logn, "parts": parts, "variant": v, "ms": round(ms, 3),
                                  "ns_per_krow": round(ms * 1e6 / (n / 1000), 2), "ok": ok}), flush=True)
            del ws
        del key, val, pairs
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
