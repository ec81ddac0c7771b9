"""Why is bench.py's sum slower than tools/sum_size_probe.py for the same bytes?"""
import sys, types, torch
sys.path.insert(0, ".")
import bench
from dpu_olap_b200.ops import Context
from dpu_olap_b200.generator import RandomArrayGenerator
D = bench.Dist(1)
ctx = Context(0)
args = types.SimpleNamespace(sf=2048, steps=10, warmup=3)
print("bench_sum cold:", {k: round(v, 3) if isinstance(v, float) else v for k, v in bench.bench_sum(ctx, D, args).items() if k in ("ms_per_step", "achieved_gbs")}, flush=True)
g = RandomArrayGenerator(ctx, 42)
col = g.batches_dev(2048, bench.SUM_BATCH)
print("col ptr % 256:", col.data_ptr() % 256, col.dtype, col.shape, col.is_contiguous())
out = torch.empty(1, dtype=torch.int64, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: ctx.sum_dev(col, out=out)); print("generated column, out=out :", round(ms, 3), "ms", round(col.numel()*4/ms/1e6, 1), "GB/s")
ms = t(lambda: ctx.sum_dev(col)); print("generated column, out=None:", round(ms, 3), "ms", round(col.numel()*4/ms/1e6, 1), "GB/s")
r = torch.empty_like(col); r.random_()
ms = t(lambda: ctx.sum_dev(r, out=out)); print("torch random column       :", round(ms, 3), "ms", round(col.numel()*4/ms/1e6, 1), "GB/s")
r.copy_(col)
ms = t(lambda: ctx.sum_dev(r, out=out)); print("copy of generated column  :", round(ms, 3), "ms", round(col.numel()*4/ms/1e6, 1), "GB/s")
