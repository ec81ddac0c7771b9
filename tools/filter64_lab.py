#!/usr/bin/env python
"""filter64_lab.py — time the 64-bit filter kernels on one GPU (tuning aid; csrc/filter64.cu).

    python -m dpu_olap_b200.build --lab
    B200OLAP_LIB=dpu_olap_b200/libb200olap_lab.so python tools/filter64_lab.py [--rows-log2 30]

Sweeps the lab switch B2_LAB_F64="prefetch distance,CTAs per SM" of the single-pass kernel (lab build
only) and the two-pass kernel; every run is checked against torch; prints ms and algorithmic GB/s
(8 B per row + 8 B per selected row).
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import numpy as np
    import torch

    from dpu_olap_b200._lib import TUNE_FILTER64_KERNEL
    from dpu_olap_b200.ops import Context
    p = argparse.ArgumentParser()
    p.add_argument("--rows-log2", type=int, default=30)
    p.add_argument("--reps", type=int, default=10)
    p.add_argument("--settings", default="1,0;0,0;2,0;3,0;1,4;1,3;2,4;1,2")
    p.add_argument("--thresholds-log2", default="62,60,63")
    args = p.parse_args()
    ctx = Context(0)
    n = 1 << args.rows_log2
    gen = torch.Generator(device="cuda").manual_seed(1)
    col = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=gen)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    end = torch.empty(1, dtype=torch.int64, device="cuda")
    total = torch.empty(1, dtype=torch.int64, device="cuda")
    ws = torch.empty(int(ctx._lib.b2_filter_64_ws_bytes(n)) + 256, dtype=torch.uint8, device="cuda")

    def run(thr):
        ctx.filter64_dev(col, np.uint64, thr, out=out, batch_end=end, total=total, ws=ws)

    def timed(thr):
        for _ in range(3):
            run(thr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run(thr)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    for tl in [int(t) for t in args.thresholds_log2.split(",")]:
        thr = 1 << tl
        keep = (col >= 0) & (col < thr) if tl < 63 else col >= 0   # unsigned v < 2^tl on the bit patterns
        ref = col[keep]
        k = ref.numel()
        del keep
        for name, kern, env in ([("two_pass", 1, None)] +
                                [(f"single pf={s.split(',')[0]} ctas={s.split(',')[1]}", 0, s)
                                 for s in args.settings.split(";")]):
            ctx.set_tunable(TUNE_FILTER64_KERNEL, kern)
            if env is None:
                os.environ.pop("B2_LAB_F64", None)
            else:
                os.environ["B2_LAB_F64"] = env
            out.zero_()
            ms = timed(thr)
            ok = int(total.item()) == k and torch.equal(out[:k], ref)
            print(f"thr=2^{tl} sel={k / n:.3f} {name:28s} {ms:8.3f} ms {(8 * n + 8 * k) / ms / 1e6:8.1f} GB/s "
                  f"{'ok' if ok else 'MISMATCH'}", flush=True)
        del ref
    ctx.close()


if __name__ == "__main__":
    main()
