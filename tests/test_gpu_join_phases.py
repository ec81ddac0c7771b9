"""Join phase timers through the ABI (b2_join_trace / b2_join_last_phases) — the reference's JoinDpu
keeps the timers "build", "probe", "take", "partitionKernel", ... (host/join/join_dpu.cc:146-148) and
join_benchmark.cc reports them. Here they are CUDA-event intervals at the phase boundaries of the
join, summed per phase."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(n, seed=3):
    rng = np.random.default_rng(seed)
    pk = rng.permutation(n).astype(np.uint32)
    x = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    fk = rng.integers(0, n, size=n, dtype=np.uint32)
    y = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    dev = lambda a: torch.from_numpy(a.view(np.int32)).cuda()
    return fk, y, pk, x, [dev(a) for a in (fk, y, pk, x)]


def test_trace_off_by_default_and_resets(ctx):

    _, _, _, _, d = _inputs(1 << 16)
    ctx.join_dev(*d)
    torch.cuda.synchronize()
    p = ctx.join_last_phases()
    assert p["intervals"] == 0 and p["probe_ms"] == 0.0
    ctx.join_trace(True)
    try:
        out = ctx.join_dev(*d)
        p = ctx.join_last_phases()          # waits for the join's events itself
        assert int(out[3].item()) == 1 << 16
        assert p["intervals"] == 3          # build side's passes | probe side's passes | probe
        assert p["partition_build_ms"] > 0 and p["partition_probe_ms"] > 0 and p["probe_ms"] > 0
        assert p["take_ms"] == 0.0
        # the next join starts a new trace: same number of intervals, not twice as many
        ctx.join_dev(*d)
        assert ctx.join_last_phases()["intervals"] == 3
    finally:
        ctx.join_trace(False)
    assert ctx.join_last_phases()["intervals"] == 0


def test_phases_add_up_to_the_join(ctx):
    """At a size where kernels dominate launch gaps the three phases account for the join's duration."""
    n = 1 << 23
    _, _, _, _, d = _inputs(n, seed=5)
    ws = torch.empty(ctx.join_ws_bytes(n, n) + 256, dtype=torch.uint8, device="cuda")
    outs = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3)]
    ctx.join_trace(True)
    try:
        ctx.join_dev(*d, ws=ws, outs=outs)   # warm-up: function attributes, first-use costs
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.join_dev(*d, ws=ws, outs=outs)
        e1.record()
        torch.cuda.synchronize()
        p = ctx.join_last_phases()
    finally:
        ctx.join_trace(False)
    total = e0.elapsed_time(e1)
    parts = p["partition_build_ms"] + p["partition_probe_ms"] + p["probe_ms"]
    assert 0.7 * total <= parts <= 1.02 * total, (total, p)


def test_sliced_join_sums_its_slices(ctx):
    n = 1 << 21
    fk, y, pk, x, d = _inputs(n, seed=4)
    full, small = ctx.join_ws_bytes(n, n), ctx.join_min_ws_bytes(n, n)
    ws = torch.empty((full + small) // 3 + 256, dtype=torch.uint8, device="cuda")
    ctx.join_trace(True)
    try:
        out = ctx.join_dev(*d, ws=ws)
        p = ctx.join_last_phases()
    finally:
        ctx.join_trace(False)
    assert int(out[3].item()) == n
    assert p["intervals"] >= 6 and p["intervals"] % 3 == 0   # 3 intervals per hash-space slice


def test_join_gpu_timers_carry_the_reference_names(ctx):
    from dpu_olap_b200 import ops
    n, nb = 1 << 18, 4
    fk, y, pk, x, _ = _inputs(n, seed=9)
    sl = lambda a, b: a[b * (n // nb):(b + 1) * (n // nb)]
    left = [{"fk": sl(fk, b), "y": sl(y, b)} for b in range(nb)]
    right = [{"pk": sl(pk, b), "x": sl(x, b)} for b in range(nb)]
    j = ops.JoinGpu(ctx, left, right)
    j.Prepare()
    out = j.Run()
    assert out["fk"].size == n
    t = j.Timers()
    for name in ("copy-to-dpu", "dpu-work", "copy-from-dpu", "partitionKernel", "probe", "take"):
        assert name in t
    assert t["partitionKernel"] > 0 and t["probe"] > 0 and t["take"] == 0.0
    assert abs(t["partitionKernel"] - t["partitionKernel-build-side"] - t["partitionKernel-probe-side"]) < 1e-9
    # the phases lie inside the device-work interval of the call
    assert t["partitionKernel-probe-side"] + t["probe"] <= t["dpu-work"] * 1.02 + 0.05
    # several payload columns: row numbers travel with the keys, the take kernel gathers the payloads
    left2 = [{**b, "y2": b["y"]} for b in left]
    j2 = ops.JoinGpu(ctx, left2, right)
    j2.Prepare()
    out2 = j2.Run()
    assert out2["fk"].size == n and j2.Timers()["take"] > 0
    assert j.RunAggregate()["rows"] == n and j.Timers()["probe"] > 0

