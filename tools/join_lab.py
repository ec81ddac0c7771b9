#!/usr/bin/env python
"""join_lab.py — time the single-GPU join for every scatter-kernel shape (tuning aid).

    python tools/join_lab.py [--sf 64,512] [--configs smem:0,smem:1,smem:2,smem:3]

`smem:v[:b[:p[:t[:d]]]]` = shared-memory tables after two radix passes, scatter-kernel shape v, whole-sector
scatter kernel from a fan-out of 2^b (default 8), next-tile prefetch p = 0 / 1 (default 1), sector-kernel
variant t (default 3 = bulk flush), perfect-hash probe path from d build rows per partition (0 = off).
Settings are sticky from one config to the next: spell every field out when comparing. (Commit 788c869
also had `l2:G:F`, the L2-resident table experiment described in profiles/r1_join_l2.md.)
"""
import argparse, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    from dpu_olap_b200.ops import Context
    p = argparse.ArgumentParser()
    p.add_argument("--sf", default="64,512")
    p.add_argument("--configs", default="smem:0,smem:1,smem:2,smem:3")
    p.add_argument("--reps", type=int, default=5)
    a = p.parse_args()
    ctx = Context(0)
    B = 2 << 20
    for sf in [int(x) for x in a.sf.split(",")]:
        g = RandomArrayGenerator(ctx, 42)
        x = g.batches_dev(sf, B)
        pk = g.index_column_dev(sf, B)
        y = g.batches_dev(sf, B)
        fk = g.foreign_key_dev(B, sf, B)
        n = sf * B
        outs = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3)]
        rows = torch.empty(1, dtype=torch.int64, device="cuda")
        fk_sum = int(fk.to(torch.int64).sum())
        for cfg in a.configs.split(","):
            f = cfg.split(":")
            assert f[0] == "smem", cfg
            ctx.set_tunable(2, int(f[1]) if len(f) > 1 else 0)   # B2_TUNE_SCATTER_SHAPE
            if len(f) > 2:  # smem:v:b = whole-sector scatter kernel from a fan-out of 2^b
                ctx.set_tunable(0, int(f[2]))                    # B2_TUNE_SCATTER_SECTORS_MIN_BITS
            if len(f) > 3:  # smem:v:b:p = next-tile prefetch in the scatter kernels off / on
                ctx.set_tunable(1, int(f[3]))                    # B2_TUNE_SCATTER_PREFETCH
            if len(f) > 4:  # smem:v:b:p:t = whole-sector kernel over 16384-row tiles, one CTA per SM
                ctx.set_tunable(4, int(f[4]))                    # B2_TUNE_SCATTER_SECTOR_TILE
            if len(f) > 5:  # smem:v:b:p:t:d = perfect-hash probe path from d build rows per partition (0 = off)
                ctx.set_tunable(5, int(f[5]))                    # B2_TUNE_JOIN_DIRECT_MIN_ROWS
            ws = torch.empty(ctx.join_ws_bytes(n, n) + 256, dtype=torch.uint8, device="cuda")
            step = lambda: ctx.join_dev(fk, y, pk, x, out_capacity=n, ws=ws, outs=outs, out_rows=rows)
            for o in outs:
                o.zero_()
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
            # every probe row matches exactly once and y / x are payloads of the same row numbers
            ok = int(rows.cpu()[0]) == n and int(outs[0].to(torch.int64).sum()) == fk_sum
            ok = ok and int(outs[1].to(torch.int64).sum()) == int(y.to(torch.int64).sum())
            xs = x[outs[0].to(torch.int64)[: 1 << 20]]
            ok = ok and bool(torch.equal(xs, outs[2][: 1 << 20]))
            print(json.dumps({"sf": sf, "config": cfg, "ms": round(ms, 3), "ws_gib": round(ws.numel() / 2**30, 2),
                              "rows_per_s": n / (ms * 1e-3), "ok": ok}), flush=True)
            del ws
        del x, pk, y, fk, outs
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
