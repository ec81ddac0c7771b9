#!/usr/bin/env python
"""join_lab.py — time the single-GPU join for every scatter-kernel shape (tuning aid).

    python tools/join_lab.py [--sf 64,512] [--variants 0,1,2,3]
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
    p.add_argument("--variants", default="0,1,2,3")
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
        ws = torch.empty(ctx.join_ws_bytes(n, n) + 256, dtype=torch.uint8, device="cuda")
        for v in [int(x) for x in a.variants.split(",")]:
            assert ctx._lib.b200olap_tune_scatter_variant(v) == 0
            step = lambda: ctx.join_dev(fk, y, pk, x, out_capacity=n, ws=ws, outs=outs, out_rows=rows)
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
            ok = int(rows.cpu()[0]) == n and int(outs[0].to(torch.int64).sum()) == int(fk.to(torch.int64).sum())
            print(json.dumps({"sf": sf, "variant": v, "ms": round(ms, 3), "rows_per_s": n / (ms * 1e-3), "ok": ok}), flush=True)
        del x, pk, y, fk, outs, ws
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
