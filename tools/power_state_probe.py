"""Does sustained load change what a pure read stream achieves? sum(16 GiB) cold, right after 2 s
of back-to-back filter launches over 32 GiB, and again after idling."""
import subprocess, sys, time, torch
sys.path.insert(0, ".")
from dpu_olap_b200.ops import Context
ctx = Context(0)
col = torch.empty(32 << 28, dtype=torch.int32, device="cuda"); col.random_()
out = torch.empty_like(col)
nb, bl = (32 << 28) // 65536, 65536
end = torch.empty(nb, dtype=torch.int64, device="cuda"); tot = torch.empty(1, dtype=torch.int64, device="cuda")
ws = torch.empty(ctx.filter_ws_bytes(nb, bl), dtype=torch.uint8, device="cuda")
def smi():
    q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown"
    return subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
def time_sum(tag):
    c = col[: 16 << 28]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.sum_dev(c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{tag:28s} sum 16 GiB: {ms:6.3f} ms {c.numel()*4/ms/1e6:7.1f} GB/s | {smi()}", flush=True)
def time_filter(tag, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ctx.filter_dev(col, nb, bl, 1 << 30, out=out, batch_end=end, total=tot, ws=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag:28s} filter 32 GiB: {ms:6.3f} ms {(col.numel()*5)/ms/1e6:7.1f} GB/s | {smi()}", flush=True)
for _ in range(3): ctx.sum_dev(col[: 16 << 28])
time_sum("cold")
time_filter("filter burst (3 launches)", 3)
time_filter("filter sustained (250)", 250)
time_sum("right after sustained load")
time_filter("filter again (10)", 10)
time.sleep(3)
time_sum("after 3 s idle")
time_filter("filter after idle (3)", 3)
