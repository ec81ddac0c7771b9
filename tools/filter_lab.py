#!/usr/bin/env python
"""filter_lab.py — time every compiled shape of the filter kernel on one GPU (tuning aid).

    python tools/filter_lab.py [--sf 256] [--variants 0,1,2,3] [--reps 10]

For each variant (b200olap_tune_filter_variant) and each threshold of the selectivity sweep it checks
the result against torch (count, per-batch ends, and full content on a slice) and prints the
achieved algorithmic HBM GB/s = (4 B/row + 4 B/selected row) / CUDA-event time.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from dpu_olap_b200.generator import RandomArrayGenerator
    from dpu_olap_b200.ops import Context
    p = argparse.ArgumentParser()
    p.add_argument("--sf", type=int, default=256)
    p.add_argument("--variants", default="0,1,2,3")
    p.add_argument("--reps", type=int, default=10)
    p.add_argument("--thresholds", default="1073741824,42949673,2147483648,4294967295")
    args = p.parse_args()
    ctx = Context(0)
    nb, bl = args.sf << 7, 65536
    n = nb * bl
    col = RandomArrayGenerator(ctx, 42).batches_dev(nb, bl)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    end = torch.empty(nb, dtype=torch.int64, device="cuda")
    total = torch.empty(1, dtype=torch.int64, device="cuda")
    ws = torch.empty(ctx.filter_ws_bytes(nb, bl), dtype=torch.uint8, device="cuda")
    flip = torch.tensor(-2**31, dtype=torch.int32, device="cuda")
    results = []
    for variant in [int(v) for v in args.variants.split(",")]:
        ctx.set_tunable(3, variant & 0xff)  # B2_TUNE_FILTER_VARIANT
        if variant >> 8:  # debug bits (no prefix / no TMA / dynamic tickets) exist in the lab build only:
            # python -m dpu_olap_b200.build --lab, then run with B200OLAP_LIB=dpu_olap_b200/libb200olap_lab.so
            rc = ctx._lib.b200olap_lab_filter_debug(variant >> 8)
            assert rc == 0, rc
        for thr in [int(t) for t in args.thresholds.split(",")]:
            def step():
                ctx.filter_dev(col, nb, bl, thr, out=out, batch_end=end, total=total, ws=ws)
            out.zero_()
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            sel = int(total.cpu()[0])
            # checks: count, batch ends, content of the first and last 2^26 rows' worth of output
            tflip = thr - 2**31
            ok = True
            if not (variant >> 8) & 1:
                cnt = 0
                for c in col.split(1 << 28):
                    cnt += int(((c ^ flip) < tflip).sum())
                ok = cnt == sel
                m = min(n, 1 << 26)
                head = col[:m]
                exp = head[(head ^ flip) < tflip]
                ok = ok and torch.equal(exp, out[: exp.numel()])
                per_batch = ((col[:m] ^ flip) < tflip).view(-1, bl).sum(1).cumsum(0)
                ok = ok and torch.equal(per_batch, end[: m // bl])
                tail = col[n - m:]
                expt = tail[(tail ^ flip) < tflip]
                ok = ok and torch.equal(expt, out[sel - expt.numel(): sel])
            else:
                sel = int(n * min(1.0, thr / 2**32))
            gbs = (4 * n + 4 * sel) / (ms * 1e-3) / 1e9
            r = {"variant": variant, "thr": thr, "sel": round(sel / n, 4), "ms": round(ms, 4),
                 "gbs": round(gbs, 1), "rows_per_s": n / (ms * 1e-3), "ok": bool(ok)}
            print(json.dumps(r), flush=True)
            results.append(r)
    bad = [r for r in results if not r["ok"]]
    ctx.close()
    if bad:
        raise SystemExit(f"{len(bad)} configurations produced wrong results")


if __name__ == "__main__":
    main()
