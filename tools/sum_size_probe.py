"""How does a pure read stream (sum kernel) scale with the footprint of the column?"""
import sys, torch
sys.path.insert(0, ".")
from dpu_olap_b200.ops import Context
ctx = Context(0)
big = torch.empty(40 << 28, dtype=torch.int32, device="cuda")  # 40 GiB
big.random_()
for gib in (1, 2, 4, 8, 12, 16, 24, 32, 40):
    col = big[: gib << 28]
    for _ in range(3):
        ctx.sum_dev(col)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, 64 // gib)
    e0.record()
    for _ in range(reps):
        ctx.sum_dev(col)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{gib:3d} GiB  {ms:8.3f} ms  {col.numel() * 4 / ms / 1e6:8.1f} GB/s", flush=True)
# same 8 GiB, but at different offsets inside the 40 GiB allocation
for off in (0, 16, 31):
    col = big[off << 28: (off + 8) << 28]
    for _ in range(3): ctx.sum_dev(col)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): ctx.sum_dev(col)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(f"8 GiB at +{off} GiB: {col.numel() * 4 / ms / 1e6:8.1f} GB/s", flush=True)
