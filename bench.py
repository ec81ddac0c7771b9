#!/usr/bin/env python
"""bench.py — rows/s of the columnar operator hot path at SF=2048 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sf 2048] [--ops filter,sweep,sum,take,join,nullable,joinsum]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # the reference's Arrow Acero CPU path on the host cores

Headline (BASELINE.json configs[1]): Filter `v < 2^30` over SF*128 batches x 65536 uint32 rows
(2^34 rows = 64 GiB at SF=2048), inputs resident in HBM, generated on the device bit-identically to
RandomArrayGenerator(42). One "step" = one pass of the operator over the whole column. The column
is row-range sharded over the N ranks with no data-path collective (total work fixed: strong
scaling); value = total rows / max-over-ranks device time.

One JSON line on stdout (rank 0). Besides the driver's keys it carries
  roofline      achieved HBM GB/s of the filter kernel vs the measured peak
  e2e           same metric through the host-buffer C ABI (b2_filter_lt_u32_host_into):
                pinned host batches in, host result out, copies inside the timed region
  cpu_baseline  Arrow Acero (the reference's CPU engine) on this box's host cores, bounded sample
  ops           the other operators of the path (sum, take, join), the selectivity sweep, the
                nullable variants and the fused join->aggregate pipelines, each timed over the same
                K steps and checked against torch on the device; `--ops ...,wide` adds the
                64-bit aggregates and take (opt-in, not in the default run)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FILTER_BATCH = 64 << 10    # filter_benchmark.cc:142,153  (SF<<7 batches of 64 Ki rows)
SUM_BATCH = 2 << 20        # aggr_benchmark.cc:132-138,148-150
TAKE_BATCH = 4 << 20       # take_benchmark.cc:157-159
TAKE_IDX = TAKE_BATCH >> 3
JOIN_BATCH = 2 << 20       # join_benchmark.cc:168-176


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--sf", type=int, default=int(os.environ.get("SF", 2048)))
    p.add_argument("--ops", default="filter,sweep,sum,take,join,nullable,joinsum")
    p.add_argument("--e2e-sf", type=int, default=int(os.environ.get("E2E_SF", 64)),
                   help="scale factor of the host-buffer (e2e) leg; bounded by host RAM and PCIe time")
    p.add_argument("--cpu-sf", type=int, default=int(os.environ.get("CPU_SF", 64)),
                   help="scale factor of the CPU (Arrow Acero) sample: BASELINE.json configs[0]")
    p.add_argument("--cpu-seconds", type=float, default=15.0)
    p.add_argument("--join-exchange", default="p2p", choices=["p2p", "nccl"],
                   help="multi-GPU join: fused peer-memory shuffle (default) or routing + NCCL all-to-all")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    return p.parse_args()


def measured_peaks() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# distributed plumbing (torch.distributed over NCCL; one process per GPU)
# ------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, want_gpus: int):
        import torch
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if self.world != want_gpus:
            if self.world == 1 and want_gpus > 1:
                raise SystemExit(f"--gpus {want_gpus} needs torchrun: python -m torch.distributed.run "
                                 f"--nnodes=1 --nproc-per-node {want_gpus} --master-addr 127.0.0.1 "
                                 f"--master-port 29500 bench.py --gpus {want_gpus}")
            raise SystemExit(f"WORLD_SIZE={self.world} but --gpus {want_gpus}")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        import torch
        if self.dist:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_float(self, v: float) -> float:
        import torch
        if not self.dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_int(self, v: int) -> int:
        import torch
        if not self.dist:
            return v
        t = torch.tensor([v], dtype=torch.int64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def close(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_steps(D: Dist, fn, steps: int, warmup: int) -> float:
    """W untimed + exactly K timed steps, barrier + synchronize on both sides, CUDA events on the
    launching (current) stream, max over ranks. Returns ms per step."""
    import torch
    for _ in range(warmup):
        fn()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    D.barrier()
    return D.max_float(ms) / steps


def shard(nbatches: int, D: Dist) -> tuple[int, int]:
    per = nbatches // D.world
    if per * D.world != nbatches:
        raise SystemExit(f"{nbatches} batches do not split over {D.world} ranks")
    return D.rank * per, per


def free_all():
    """Release this operator's columns before the next one allocates its own. Giving tens of GiB back
    to the driver is not free: for a while afterwards (memory being scrubbed, power state of the
    sustained run before) the next kernels ran ~20 % slower when they started immediately
    (tools/sum_bench_probe.py: 7.2 TB/s cold vs 5.8 TB/s straight after the filter sweep). BENCH_SETTLE_S
    (default 1 s) lets the device settle between OPERATORS; it is outside every timed region."""
    import gc
    import torch
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    settle = float(os.environ.get("BENCH_SETTLE_S", "1.0"))
    if settle > 0:
        time.sleep(settle)



# ------------------------------------------------------------------------------------------------
# full-size content checks (torch on the device; independent of the library's kernels)
# ------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    """A 64-bit pattern as the signed Python int torch's int64 arithmetic wants."""
    v &= _M64
    return v - (1 << 64) if v >> 63 else v


def _lsr(z, k: int):
    return (z >> k) & ((1 << (64 - k)) - 1)  # logical shift right on int64


def _mix64(z):
    """oracle/olap_oracle.c:mix64 on int64 tensors (products wrap mod 2^64 like the C code's)."""
    z = z ^ _lsr(z, 33)
    z = z * _s64(0xff51afd7ed558ccd)
    z = z ^ _lsr(z, 33)
    z = z * _s64(0xc4ceb9fe1a85ec53)
    return z ^ _lsr(z, 33)


def triple_checksum_torch(a, b, c, chunk: int = 1 << 26) -> int:
    """Order-independent 64-bit checksum of the (a, b, c) row multiset: the torch restatement of
    orc_triple_checksum (oracle/olap_oracle.c:199-211; tests/test_bench_checks.py pins the two to
    each other). a, b, c: int32 / uint32-as-int32 tensors of equal length, any device."""
    import torch
    total = 0
    n = a.numel()
    for s0 in range(0, n, chunk):
        s1 = min(n, s0 + chunk)
        aa = a[s0:s1].to(torch.int64) & 0xFFFFFFFF
        bb = b[s0:s1].to(torch.int64) & 0xFFFFFFFF
        cc = c[s0:s1].to(torch.int64) & 0xFFFFFFFF
        z = _mix64(((aa << 32) | bb) ^ _mix64(cc + _s64(0x9e3779b97f4a7c15)))
        total = (total + int(z.sum())) & _M64
    return total


def check_filter_content(col, nb: int, out, end, thr: int, what: str) -> int:
    """Every selected row and every batch boundary of a filter result, streamed in 2^28-row chunks:
    torch's boolean selection of chunk k must equal out[pos .. pos + k) and the running per-batch
    counts must equal batch_end. Returns the number of selected rows."""
    import torch
    flip = torch.tensor(-2**31, dtype=torch.int32, device=col.device)
    tflip = int(thr) - 2**31  # unsigned compare via sign flip
    pos = 0
    cb = max(1, (1 << 28) // FILTER_BATCH)
    for b0 in range(0, nb, cb):
        b1 = min(nb, b0 + cb)
        c = col[b0 * FILTER_BATCH:b1 * FILTER_BATCH]
        m = (c ^ flip) < tflip
        exp = c[m]
        k = exp.numel()
        if not torch.equal(exp, out[pos:pos + k]):
            raise SystemExit(f"{what}: selected rows of batches [{b0}, {b1}) differ from torch's selection")
        cum = m.view(b1 - b0, FILTER_BATCH).sum(1).cumsum(0) + pos
        if not torch.equal(cum, end[b0:b1]):
            raise SystemExit(f"{what}: batch_end of batches [{b0}, {b1}) differs from torch's counts")
        pos += k
        del c, m, exp, cum
    return pos


def ncu_traffic(op: str):
    """DRAM bytes per row of an operator's kernels from the ncu capture committed for this round
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py from the ncu CSVs named in it).
    None when there is no capture — never a constant made up here."""
    f = ROOT / "profiles" / "r2_traffic.json"
    try:
        return json.loads(f.read_text()).get(op)
    except Exception:
        return None


def filter_traffic(fres: dict, world: int) -> dict:
    tr = ncu_traffic("filter")
    if not tr or tr.get("dram_bytes_per_algorithmic_byte") is None:
        return {"traffic": None, "traffic_source": "no ncu capture committed for this round"}
    return {"traffic": tr["dram_bytes_per_algorithmic_byte"] * fres["algorithmic_bytes"] / world,
            "traffic_per_algorithmic_byte": tr["dram_bytes_per_algorithmic_byte"],
            "traffic_source": tr.get("source", "profiles/r2_traffic.json")}


def roofline_block(op: str, algorithmic_bytes: float, ms: float, peak: float, peak_src: str, rows: float,
                   world: int, note: str, kernel: str) -> dict:
    """Per-GPU roofline of one operator step: algorithmic bytes / CUDA-event time against the
    measured copy bandwidth AND the nominal 8 TB/s SURVEY.md section 8(d) names."""
    achieved = algorithmic_bytes / world / (ms * 1e-3) / 1e9
    tr = ncu_traffic(op)
    blk = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "frac_of_nominal_8000": achieved / 8000.0, "peak_source": peak_src, "kernel": kernel,
           "algorithmic_bytes_per_launch": algorithmic_bytes / world,
           "algorithmic_bytes_per_row": algorithmic_bytes / rows, "note": note,
           "traffic": None, "traffic_source": "no ncu capture committed for this operator"}
    if tr and tr.get("dram_bytes_per_row") is not None:
        blk["traffic"] = tr["dram_bytes_per_row"] * rows / world
        blk["traffic_bytes_per_row"] = tr["dram_bytes_per_row"]
        blk["traffic_source"] = tr.get("source", "profiles/r2_traffic.json")
    return blk


# ------------------------------------------------------------------------------------------------
# operators
# ------------------------------------------------------------------------------------------------
def bench_filter(ctx, D, args, thresholds):
    """Returns {thr: dict} for every threshold, timed on the same resident column."""
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    nb_total = args.sf << 7
    first, nb = shard(nb_total, D)
    g = RandomArrayGenerator(ctx, 42)
    col = g.batches_dev(nb_total, FILTER_BATCH, take=(first, nb))
    n = nb * FILTER_BATCH
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    end = torch.empty(nb, dtype=torch.int64, device="cuda")
    total = torch.empty(1, dtype=torch.int64, device="cuda")
    ws = torch.empty(ctx.filter_ws_bytes(nb, FILTER_BATCH), dtype=torch.uint8, device="cuda")
    res = {}
    for thr in thresholds:
        def step():
            ctx.filter_dev(col, nb, FILTER_BATCH, thr, out=out, batch_end=end, total=total, ws=ws)
        l0 = ctx.launches
        ms = timed_steps(D, step, args.steps, args.warmup)
        launches = (ctx.launches - l0) // (args.steps + args.warmup)
        sel_local = int(total.cpu()[0])
        # full-size CONTENT check: every output row and every batch boundary against torch
        cnt = check_filter_content(col, nb, out, end, thr, f"filter self-check (rank {D.rank}, thr {thr})")
        if cnt != sel_local:
            raise SystemExit(f"filter self-check failed on rank {D.rank}: {cnt} vs {sel_local}")
        sel = D.sum_int(sel_local)
        rows = nb_total * FILTER_BATCH
        res[thr] = {"ms_per_step": ms, "rows": rows, "selected": sel, "rows_per_s": rows / (ms * 1e-3),
                    "algorithmic_bytes": 4 * rows + 4 * sel, "launches_per_step": launches,
                    "rows_per_rank": n, "self_check": "content"}
    del col, out, end, total, ws
    free_all()
    return res


def bench_sum(ctx, D, args):
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    nb_total = args.sf
    first, nb = shard(nb_total, D) if nb_total >= D.world else (0, 0)
    if nb == 0:
        return None
    g = RandomArrayGenerator(ctx, 42)
    col = g.batches_dev(nb_total, SUM_BATCH, take=(first, nb))
    out = torch.empty(1, dtype=torch.int64, device="cuda")
    ms = timed_steps(D, lambda: ctx.sum_dev(col, out=out), args.steps, args.warmup)
    got = int(out.cpu().numpy().view("uint64")[0])
    chk = 0
    for c in col.split(1 << 28):
        chk += int((c.to(torch.int64) & 0xFFFFFFFF).sum())
    if got != chk % (1 << 64):
        raise SystemExit(f"sum self-check failed on rank {D.rank}")
    rows = nb_total * SUM_BATCH
    del col
    free_all()
    return {"ms_per_step": ms, "rows": rows, "rows_per_s": rows / (ms * 1e-3), "self_check": "content",
            "achieved_gbs": 4 * rows / D.world / (ms * 1e-3) / 1e9, "algorithmic_bytes_per_row": 4,
            "sum_rank0": got}


def bench_take(ctx, D, args):
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    nb_total = args.sf
    first, nb = shard(nb_total, D) if nb_total >= D.world else (0, 0)
    if nb == 0:
        return None
    g = RandomArrayGenerator(ctx, 42)
    vals = g.batches_dev(nb_total, TAKE_BATCH, take=(first, nb))          # take_benchmark.cc:86-95:
    idx = g.batches_dev(nb_total, TAKE_IDX, 0, TAKE_BATCH - 1, take=(first, nb))  # all v, then all i
    out = torch.empty(nb * TAKE_IDX, dtype=torch.int32, device="cuda")
    ms = timed_steps(D, lambda: ctx.take_dev(vals, TAKE_BATCH, idx, TAKE_IDX, nb, out=out),
                     args.steps, args.warmup)
    # full-size content check: every batch against torch's gather (64 batches per torch call)
    for b0 in range(0, nb, 64):
        b1 = min(nb, b0 + 64)
        gi = idx[b0 * TAKE_IDX:b1 * TAKE_IDX].to(torch.int64).view(b1 - b0, TAKE_IDX)
        gi = gi + torch.arange(b1 - b0, device="cuda", dtype=torch.int64)[:, None] * TAKE_BATCH
        ref = vals[b0 * TAKE_BATCH:b1 * TAKE_BATCH][gi.view(-1)]
        if not torch.equal(ref, out[b0 * TAKE_IDX:b1 * TAKE_IDX]):
            raise SystemExit(f"take self-check failed on rank {D.rank}, batches [{b0}, {b1})")
        del gi, ref
    nidx = nb_total * TAKE_IDX
    del vals, idx, out
    free_all()
    return {"ms_per_step": ms, "indices": nidx, "value_rows": nb_total * TAKE_BATCH, "self_check": "content",
            "rows_per_s": nidx / (ms * 1e-3), "algorithmic_bytes_per_index": 12,
            "achieved_gbs": 12 * nidx / D.world / (ms * 1e-3) / 1e9,
            # what HBM really moves: at 1 index per 8 values nearly every 64 B unit of the batch
            # window is touched (profiles/r1_sum_take.md), so the window itself is the traffic
            "window_gbs": (4 * nb_total * TAKE_BATCH + 8 * nidx) / D.world / (ms * 1e-3) / 1e9,
            "reference_convention_rows_per_s": nb_total * TAKE_BATCH / (ms * 1e-3)}


def bench_wide(ctx, D, args):
    """Opt-in (`--ops ...,wide`; not part of the default run): the 64-bit operators of SURVEY.md
    section 8f-3 on resident columns — one-pass aggregates over a uint64 column and take over 64-bit
    values with 32-bit indices. Bounded at SF=256; results checked against torch on the device."""
    import numpy as np
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    from dpu_olap_b200.ops import decode_aggr
    sf = min(args.sf, 256)
    nb_total = sf << 7
    first, nb = shard(nb_total, D)
    g = RandomArrayGenerator(ctx, 42)
    words = g.batches_dev(nb_total, FILTER_BATCH, take=(first, nb))  # two 32-bit draws per 64-bit value
    col = words.view(torch.int64)
    n = col.numel()
    rows = nb_total * FILTER_BATCH // 2
    res = {"sf": sf}
    aout = torch.empty(4, dtype=torch.int64, device="cuda")
    ms = timed_steps(D, lambda: ctx.aggr_dev(col, None, out=aout, dtype=np.uint64), args.steps, args.warmup)
    agg = decode_aggr(aout, np.uint64)
    flip = torch.tensor(-2**63, dtype=torch.int64, device="cuda")
    ref_sum = int(col.sum()) % (1 << 64)  # torch's int64 sum wraps like Arrow's unchecked sum
    ref_min = (int((col ^ flip).min()) + 2**63) % (1 << 64)
    ref_max = (int((col ^ flip).max()) + 2**63) % (1 << 64)
    if (agg["sum"], agg["count"], agg["min"], agg["max"]) != (ref_sum, n, ref_min, ref_max):
        raise SystemExit(f"64-bit aggregate self-check failed on rank {D.rank}")
    res["aggregates_u64"] = {"ms_per_step": ms, "rows": rows, "rows_per_s": rows / (ms * 1e-3),
                             "algorithmic_bytes_per_row": 8,
                             "achieved_gbs": 8 * rows / D.world / (ms * 1e-3) / 1e9, "result_rank0": agg}
    # ---- filter over the same uint64 column: v < 2^62 (25 % selected), batches of 32768 rows ----
    from dpu_olap_b200._lib import TUNE_FILTER64_KERNEL
    thr = 1 << 62
    bl = FILTER_BATCH // 2
    fout = torch.empty(n, dtype=torch.int64, device="cuda")
    fend = torch.empty(nb, dtype=torch.int64, device="cuda")
    ftot = torch.empty(1, dtype=torch.int64, device="cuda")
    fws = torch.empty(int(ctx._lib.b2_filter_64_ws_bytes(n)) + 256, dtype=torch.uint8, device="cuda")
    fres = {}
    for kern, name in ((1, "two_pass"), (0, "single_pass")):   # the default kernel last: its output is checked
        ctx.set_tunable(TUNE_FILTER64_KERNEL, kern)
        fres[name] = timed_steps(D, lambda: ctx.filter64_dev(col, np.uint64, thr, nbatches=nb, batch_len=bl, out=fout,
                                                             batch_end=fend, total=ftot, ws=fws),
                                 args.steps, args.warmup)
    sel = 0
    chunk = 1 << 27
    for r0 in range(0, n, chunk):   # content check, streamed: unsigned v < 2^62  <=>  0 <= signed v < 2^62
        c = col[r0:r0 + chunk]
        ref = c[(c >= 0) & (c < thr)]
        if not torch.equal(ref, fout[sel:sel + ref.numel()]):
            raise SystemExit(f"64-bit filter self-check failed on rank {D.rank}, rows [{r0}, {r0 + chunk})")
        sel += ref.numel()
        del c, ref
    if sel != int(ftot.item()) or sel != int(fend[-1].item()):
        raise SystemExit(f"64-bit filter self-check failed on rank {D.rank}: total")
    sel_all = D.sum_int(sel)
    ms = fres["single_pass"]
    res["filter_u64"] = {"ms_per_step": ms, "rows": rows, "selected": sel_all, "rows_per_s": rows / (ms * 1e-3),
                         "algorithmic_bytes_per_row": 8 + 8 * sel_all / rows, "self_check": "content",
                         "achieved_gbs": (8 * rows + 8 * sel_all) / D.world / (ms * 1e-3) / 1e9,
                         "two_pass_ms_per_step": fres["two_pass"],
                         "kernel": "filter64_single_pass_kernel (counted sums, ring of counted tiles; csrc/filter64.cu)"}
    del words, col, fout, fend, ftot, fws
    free_all()
    if sf >= D.world:
        firstb, nbt = shard(sf, D)
        vl = TAKE_BATCH // 2                                            # 2 Mi 64-bit values = 16 MiB per batch
        vals = g.batches_dev(sf, TAKE_BATCH, take=(firstb, nbt)).view(torch.int64)
        idx = g.batches_dev(sf, TAKE_IDX, 0, vl - 1, take=(firstb, nbt))
        out = torch.empty(nbt * TAKE_IDX, dtype=torch.int64, device="cuda")
        ms = timed_steps(D, lambda: ctx.take64_dev(vals, vl, idx, TAKE_IDX, nbt, out=out), args.steps, args.warmup)
        b = nbt // 2
        ref = vals[b * vl:(b + 1) * vl][idx[b * TAKE_IDX:(b + 1) * TAKE_IDX].to(torch.int64)]
        if not torch.equal(ref, out[b * TAKE_IDX:(b + 1) * TAKE_IDX]):
            raise SystemExit(f"64-bit take self-check failed on rank {D.rank}")
        nidx = sf * TAKE_IDX
        res["take_u64"] = {"ms_per_step": ms, "indices": nidx, "rows_per_s": nidx / (ms * 1e-3),
                           "algorithmic_bytes_per_index": 20,
                           "achieved_gbs": 20 * nidx / D.world / (ms * 1e-3) / 1e9}
        del vals, idx, out
        free_all()
    return res


def bench_nullable(ctx, D, args):
    """The nullable variants on resident columns with 12.5 % nulls: filter
    (valid AND v < 2^30), one-pass aggregates, take with nullable values and indices. Bounded at
    SF=256 so the default run stays short; every result is checked against torch on the device."""
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    from dpu_olap_b200.ops import decode_aggr
    sf = min(args.sf, 256)
    nb_total = sf << 7
    first, nb = shard(nb_total, D)
    g = RandomArrayGenerator(ctx, 42)
    col = g.batches_dev(nb_total, FILTER_BATCH, take=(first, nb))
    n = nb * FILTER_BATCH

    def random_bitmap(bits: int, seed: int):
        gen = torch.Generator(device="cuda").manual_seed(seed + D.rank)
        nbytes = (bits + 31) // 32 * 4 + 4
        b = torch.randint(0, 256, (3, nbytes), dtype=torch.uint8, device="cuda", generator=gen)
        return b[0] | b[1] | b[2]  # a bit is set with probability 7/8

    def expand(bitmap, lo: int, hi: int):  # rows [lo, hi) of the bitmap as bool, lo % 8 == 0
        by = bitmap[lo // 8:(hi + 7) // 8]
        sh = torch.arange(8, dtype=torch.uint8, device="cuda")
        return ((by[:, None] >> sh) & 1).bool().reshape(-1)[: hi - lo]

    valid = random_bitmap(n, 1)
    res = {"sf": sf, "null_fraction": 0.125}
    # ---- filter ----
    thr = 1 << 30
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    end = torch.empty(nb, dtype=torch.int64, device="cuda")
    total = torch.empty(1, dtype=torch.int64, device="cuda")
    ws = torch.empty(ctx.filter_ws_bytes(nb, FILTER_BATCH), dtype=torch.uint8, device="cuda")
    def fstep():
        ctx.filter_nullable_dev(col, valid, nb, FILTER_BATCH, thr, out=out, batch_end=end, total=total, ws=ws)
    ms = timed_steps(D, fstep, args.steps, args.warmup)
    sel_local = int(total.cpu()[0])
    flip = torch.tensor(-2**31, dtype=torch.int32, device="cuda")
    cnt, vsum, vcnt = 0, 0, 0
    step_rows = 1 << 26
    for lo in range(0, n, step_rows):
        hi = min(n, lo + step_rows)
        c, bits = col[lo:hi], expand(valid, lo, hi)
        cnt += int((((c ^ flip) < (thr - 2**31)) & bits).sum())
        vsum += int(((c.to(torch.int64) & 0xFFFFFFFF) * bits).sum())
        vcnt += int(bits.sum())
    if cnt != sel_local:
        raise SystemExit(f"nullable filter self-check failed on rank {D.rank}: {cnt} vs {sel_local}")
    sel = D.sum_int(sel_local)
    rows = nb_total * FILTER_BATCH
    fbytes = (4 + 1 / 8) * rows + 4 * sel
    res["filter"] = {"ms_per_step": ms, "rows": rows, "selected": sel, "rows_per_s": rows / (ms * 1e-3),
                     "algorithmic_bytes_per_row": fbytes / rows,
                     "achieved_gbs": fbytes / D.world / (ms * 1e-3) / 1e9}
    del out, end, total, ws
    # ---- aggregates ----
    aout = torch.empty(3, dtype=torch.int64, device="cuda")
    ms = timed_steps(D, lambda: ctx.aggr_dev(col, valid, out=aout), args.steps, args.warmup)
    agg = decode_aggr(aout)
    if agg["sum"] != vsum % (1 << 64) or agg["count"] != vcnt:
        raise SystemExit(f"nullable aggregate self-check failed on rank {D.rank}")
    res["aggregates"] = {"ms_per_step": ms, "rows": rows, "rows_per_s": rows / (ms * 1e-3),
                         "algorithmic_bytes_per_row": 4 + 1 / 8,
                         "achieved_gbs": (4 + 1 / 8) * rows / D.world / (ms * 1e-3) / 1e9,
                         "result_rank0": agg}
    del col, valid
    free_all()
    # ---- take ----
    if sf >= D.world:
        first, nbt = shard(sf, D)
        vals = g.batches_dev(sf, TAKE_BATCH, take=(first, nbt))
        idx = g.batches_dev(sf, TAKE_IDX, 0, TAKE_BATCH - 1, take=(first, nbt))
        vvalid, ivalid = random_bitmap(nbt * TAKE_BATCH, 2), random_bitmap(nbt * TAKE_IDX, 3)
        tout = torch.empty(nbt * TAKE_IDX, dtype=torch.int32, device="cuda")
        tbits = torch.zeros(nbt * TAKE_IDX // 8 + 8, dtype=torch.uint8, device="cuda")
        ms = timed_steps(D, lambda: ctx.take_nullable_dev(vals, vvalid, TAKE_BATCH, idx, ivalid, TAKE_IDX, nbt,
                                                          out=tout, out_valid=tbits), args.steps, args.warmup)
        b = nbt // 2
        ii = idx[b * TAKE_IDX:(b + 1) * TAKE_IDX].to(torch.int64)
        ok = expand(ivalid, b * TAKE_IDX, (b + 1) * TAKE_IDX) & \
            expand(vvalid, b * TAKE_BATCH, (b + 1) * TAKE_BATCH)[ii]
        ref = torch.where(ok, vals[b * TAKE_BATCH:(b + 1) * TAKE_BATCH][ii], torch.zeros_like(ii, dtype=torch.int32))
        if not (torch.equal(ref, tout[b * TAKE_IDX:(b + 1) * TAKE_IDX]) and
                torch.equal(ok, expand(tbits, b * TAKE_IDX, (b + 1) * TAKE_IDX))):
            raise SystemExit(f"nullable take self-check failed on rank {D.rank}")
        nidx = sf * TAKE_IDX
        res["take"] = {"ms_per_step": ms, "indices": nidx, "rows_per_s": nidx / (ms * 1e-3),
                       "algorithmic_bytes_per_index": 12 + 3 / 8,
                       "achieved_gbs": (12 + 3 / 8) * nidx / D.world / (ms * 1e-3) / 1e9}
        del vals, idx, vvalid, ivalid, tout, tbits
        free_all()
    return res


def bench_join(ctx, D, args):
    """Join at SF (2 Mi rows per batch per side). N=1: local join. N>1: both sides are routed by
    the top log2(N) hash bits (b2_shuffle_partition), exchanged with an NCCL all-to-all over
    NVLink and joined locally with hash_skip_bits = log2(N)."""
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    nb_total = args.sf
    if nb_total < D.world:
        return None
    first, nb = shard(nb_total, D)
    g = RandomArrayGenerator(ctx, 42)
    # fixture draw order (join_benchmark.cc:83-100): all x batches, then all y, then all fk
    x = g.batches_dev(nb_total, JOIN_BATCH, take=(first, nb))
    pk = g.index_column_dev(nb_total, JOIN_BATCH, take=(first, nb))
    y = g.batches_dev(nb_total, JOIN_BATCH, take=(first, nb))
    fk = g.foreign_key_dev(JOIN_BATCH, nb_total, JOIN_BATCH, take=(first, nb))
    n = nb * JOIN_BATCH
    info = {}
    # "SF=2048 and above": beyond SF=2048 the reference generator's uint32 pk counter wraps
    # (generator.cc:60,66), so every key occurs SF/2048 times and every probe row matches that many
    # build rows — a true multi-match inner join (Arrow's semantics; the DPU table would overwrite)
    mult = 1
    if nb_total * JOIN_BATCH > 1 << 32:
        if (nb_total * JOIN_BATCH) % (1 << 32):
            raise SystemExit("beyond SF=2048 the join bench needs SF to be a multiple of 2048")
        mult = (nb_total * JOIN_BATCH) >> 32
        info["matches_per_probe_row"] = mult
    if D.world == 1:
        # the three output columns are ONE allocation: dead until the probe writes them, the library
        # uses them as the temporary of the first radix pass (b200olap.h, "adjacent output columns"),
        # which is what lets SF=2048 run in one hash-space slice on one GPU
        out_all = torch.empty(3 * n * mult, dtype=torch.int32, device="cuda")
        outs = [out_all[i * n * mult:(i + 1) * n * mult] for i in range(3)]
        rows_t = torch.empty(1, dtype=torch.int64, device="cuda")
        free, _ = torch.cuda.mem_get_info()
        full = ctx.join_ws_bytes(n, n)
        one_go = ctx.join_ws_bytes_adjacent_outputs(n, n)
        margin = 1 << 30
        ws_bytes = min(full, max(free - margin, ctx.join_min_ws_bytes(n, n)))
        if one_go <= ws_bytes < full:
            ws_bytes = one_go
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
        info["workspace_gib"] = round(ws_bytes / 2**30, 2)
        info["sliced"] = ws_bytes < one_go
        info["pass1_temporary"] = "workspace" if ws_bytes >= full else ("output columns" if ws_bytes >= one_go else "workspace (sliced)")

        def step():
            ctx.join_dev(fk, y, pk, x, out_capacity=n * mult, ws=ws, outs=outs, out_rows=rows_t)
        l0 = ctx.launches
        ms = timed_steps(D, step, args.steps, args.warmup)
        info["launches_per_step"] = (ctx.launches - l0) // (args.steps + args.warmup)
        out_rows = int(rows_t.cpu().numpy().view("uint64")[0])
        o_fk, o_y, o_x = outs
        del ws, out_all
    else:
        G = D.world
        cap = n + n // 8 + 65536  # received rows: hash-uniform, 12.5 % slack
        outs = [torch.empty(cap * mult, dtype=torch.int32, device="cuda") for _ in range(3)]
        rows_t = torch.empty(1, dtype=torch.int64, device="cuda")
        exchange = args.join_exchange
        pj = None
        if exchange == "p2p":
            # fused shuffle: the routing kernel stores straight into the peers' receive buffers.
            # Symmetric memory needs peer access between all ranks; if any rank cannot set it up,
            # every rank takes the NCCL all-to-all path instead (both are GPU paths).
            from dpu_olap_b200.sharded import P2PShuffleJoin
            ok = 1
            try:
                pj = P2PShuffleJoin(ctx, D.dist, D.rank, G, n, cap, n_build_total=nb_total * JOIN_BATCH)
            except Exception as e:  # noqa: BLE001 - reported below, and agreed on by all ranks
                ok = 0
                info["p2p_unavailable"] = f"{type(e).__name__}: {e}"[:200]
            if D.sum_int(ok) != G:
                pj, exchange = None, "nccl"
        if exchange == "p2p":
            jws_bytes = ctx.join_seg_cap_ws_bytes(cap, cap, pj.nr_expected, pj.skip, pj.seg_bits)
            if jws_bytes == 0:
                raise SystemExit("segmented join unsupported at this size; use --join-exchange nccl")
            jws = torch.empty(jws_bytes + 256, dtype=torch.uint8, device="cuda")
            info["workspace_gib"] = round(jws_bytes / 2**30, 2)
            info["sliced"] = False

            def local_join(l_buf, lseg, r_buf, rseg, nr_expected, seg_bits, skip_bits, abort, phase_bits):
                ctx.join_pairs_seg_cap_dev(l_buf, lseg, r_buf, rseg, nr_expected, seg_bits, out_capacity=cap * mult,
                                           skip_bits=skip_bits, ws=jws, outs=outs, out_rows=rows_t, abort=abort,
                                           phases=phase_bits)

            def step():
                pj.step(fk, y, pk, x, local_join)
            l0 = ctx.launches
            ms = timed_steps(D, step, args.steps, args.warmup)
            info["launches_per_step"] = (ctx.launches - l0) // (args.steps + args.warmup)
            ph = {}
            pj.step(fk, y, pk, x, local_join, phases=ph)  # one extra, synchronised step: where the time goes
            info["phases_ms_rank0_serialised"] = {k: round(v, 3) for k, v in ph.items()}
            info["overlap"] = (f"timed steps: build-side scatter | barrier | probe-side scatter in {pj.shares} share(s) "
                               f"on a second stream ({pj.probe_scatter_ctas or 'all'} CTAs) UNDER the build side's fine "
                               "pass and the previous share's fine pass + probe; the serialised phase list above "
                               "comes from one extra, synchronised step")
            info["probe_shares"] = pj.shares
            info["nvlink_gbs_per_direction_serialised"] = round(
                16 * n * (G - 1) / G / (ph["scatter_nvlink"] * 1e-3) / 1e9, 1) if ph.get("scatter_nvlink") else None
            nl_r, nr_r = pj.received()
            info["shuffle_rows_received_rank0"] = [nl_r, nr_r]
            info["host_syncs_per_step"] = 0
            info["shuffle_bytes_sent_per_rank"] = int(16 * n * (G - 1) / G)  # expectation: hash-uniform
            info["shuffle"] = ("fused: b2_shuffle_p2p_scatter stores (key, payload) pairs over NVLink into the "
                               "peers' symmetric-memory receive buffers, pre-partitioned; addresses, capacity check "
                               "and received row counts stay on the device (b2_shuffle_p2p_plan_dev, "
                               "b2_join_pairs_seg_cap_dev); collectives left: one all-gather of 2 x 1025 boundaries + "
                               "one barrier (dpu_olap_b200.sharded.P2PShuffleJoin)")
            del jws, pj
        else:
            lp = torch.empty(n, dtype=torch.int64, device="cuda")
            rp = torch.empty(n, dtype=torch.int64, device="cuda")
            loff = torch.empty(G + 1, dtype=torch.int64, device="cuda")
            roff = torch.empty(G + 1, dtype=torch.int64, device="cuda")
            lrecv = torch.empty(cap, dtype=torch.int64, device="cuda")
            rrecv = torch.empty(cap, dtype=torch.int64, device="cuda")
            sws = torch.empty(int(ctx._lib.b2_shuffle_ws_bytes(n, G)) + 512, dtype=torch.uint8, device="cuda")
            free, _ = torch.cuda.mem_get_info()
            full = ctx.join_ws_bytes(cap, cap)
            jws_bytes = min(full, max(free - (3 << 30), ctx.join_min_ws_bytes(cap, cap)))
            jws = torch.empty(jws_bytes + 256, dtype=torch.uint8, device="cuda")
            info["workspace_gib"] = round(jws_bytes / 2**30, 2)
            info["sliced"] = jws_bytes < full
            from dpu_olap_b200.sharded import ShardedJoin

            def route_l(key, val):
                ctx.shuffle_partition_dev(key, val, G, pairs_out=lp, dest_off=loff, ws=sws)
                return lp, loff

            def route_r(key, val):
                ctx.shuffle_partition_dev(key, val, G, pairs_out=rp, dest_off=roff, ws=sws)
                return rp, roff

            def local_join(lr, rr, skip_bits):
                ctx.join_pairs_dev(lr, rr, out_capacity=cap * mult, skip_bits=skip_bits, ws=jws, outs=outs, out_rows=rows_t)

            sj = ShardedJoin(D.dist, D.rank, G, route_l, local_join, route_r=route_r, recv_l=lrecv, recv_r=rrecv)

            def step():
                sj.step(fk, y, pk, x)
            l0 = ctx.launches
            ms = timed_steps(D, step, args.steps, args.warmup)
            info["launches_per_step"] = (ctx.launches - l0) // (args.steps + args.warmup)
            info["shuffle_bytes_sent_per_rank"] = sj.bytes_sent()
            info["shuffle"] = "b2_shuffle_partition + NCCL all_to_all_single (NVLink), dpu_olap_b200.sharded.ShardedJoin"
            del lp, rp, lrecv, rrecv, sws, jws
        out_rows = int(rows_t.cpu().numpy().view("uint64")[0])
        o_fk, o_y, o_x = outs
    # self-check: pk is the global row index and x is drawn per pk batch, so x must equal the R.x
    # row fk points at — verified here for the rows whose pk batch this rank generated
    # (N=1: all of them), plus the row count: every fk matches exactly one pk.
    total_rows = D.sum_int(out_rows)
    if total_rows != nb_total * JOIN_BATCH * mult:
        raise SystemExit(f"join self-check failed: {total_rows} rows, expected {nb_total * JOIN_BATCH * mult}")
    if mult > 1:
        # wrapped keys: every probe row appears `mult` times in the output, so the output's fk and y
        # columns must sum to mult x the probe side's (summed over all ranks)
        def colsum(t, rows):
            acc = 0
            for s0 in range(0, rows, 1 << 26):
                acc += int((t[s0:min(s0 + (1 << 26), rows)].to(torch.int64) & 0xFFFFFFFF).sum())
            return acc
        M = 1 << 56  # per-rank residues, so that the int64 all-reduce over <= 8 ranks cannot overflow
        got = [D.sum_int(colsum(o_fk, out_rows) % M), D.sum_int(colsum(o_y, out_rows) % M)]
        exp = [D.sum_int(colsum(fk, n) * mult % M), D.sum_int(colsum(y, n) * mult % M)]
        if [g % M for g in got] != [e % M for e in exp]:
            raise SystemExit(f"join self-check failed on the multi-match sums: {got} vs {exp}")
    del pk
    free_all()
    check = "row count + column sums (wrapped keys: the matches of a probe row live on other ranks)"
    if mult == 1:
        # full-size CONTENT check, every rank's rows: the order-independent 64-bit checksum of the
        # output's (fk, y, x) multiset (torch restatement of orc_triple_checksum) against the same
        # checksum of (fk, y, R.x[fk]) computed from the INPUTS. fk of L batch b lies in pk batch b
        # (generator.cc:46-57) and pk is the global row number, so the R row a probe row matches is
        # the one this rank generated itself: x[fk - first pk of the rank]. Summed over all ranks
        # (mod 2^64) the two must agree — this covers y, x and every rank's share of the output.
        lo_pk = first * JOIN_BATCH
        exp = 0
        for s0 in range(0, n, 1 << 26):
            s1 = min(n, s0 + (1 << 26))
            kf = (fk[s0:s1].to(torch.int64) & 0xFFFFFFFF) - lo_pk
            if int(kf.min()) < 0 or int(kf.max()) >= n:
                raise SystemExit("join self-check: a foreign key points outside this rank's pk range")
            exp = (exp + triple_checksum_torch(fk[s0:s1], y[s0:s1], x[kf])) & _M64
            del kf
        got = triple_checksum_torch(o_fk[:out_rows], o_y[:out_rows], o_x[:out_rows])
        # all-reduce as 4 x 16-bit limbs so the int64 sum over <= 8 ranks cannot overflow
        def allsum64(v):
            limbs = [D.sum_int((v >> (16 * i)) & 0xFFFF) for i in range(4)]
            return sum(l << (16 * i) for i, l in enumerate(limbs)) & _M64
        got_all, exp_all = allsum64(got), allsum64(exp)
        if got_all != exp_all:
            raise SystemExit(f"join self-check failed: multiset checksum {got_all:#x} vs expected {exp_all:#x}")
        check = "content"
    del fk, y
    free_all()
    rows = nb_total * JOIN_BATCH
    res = {"ms_per_step": ms, "rows_per_side": rows, "out_rows": total_rows, "self_check": check,
           "rows_per_s": rows / (ms * 1e-3),  # probe (L) rows per second
           "items_per_s_reference_convention": 4 * rows / (ms * 1e-3),  # join_benchmark.cc:114-125
           # 8 B per build row + 8 B per probe row + 12 B per output row (mult output rows per probe row)
           "algorithmic_bytes_per_row": 16 + 12 * mult,
           "achieved_gbs_algorithmic": (16 + 12 * mult) * rows / D.world / (ms * 1e-3) / 1e9}
    res.update(info)
    del x, outs, o_fk, o_y, o_x
    free_all()
    return res


def bench_join_aggr(ctx, D, args):
    """Fused pipeline [filter L.y < 2^30 ->] join -> COUNT / SUM(y) / SUM(x) on one GPU
    (b2_join_aggr_u32_dev): same inputs as the join, nothing materialised — so at
    SF=2048 the 48 GiB of output columns are not needed and the join runs in one hash-space slice."""
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    if D.world != 1:
        return bench_join_aggr_sharded(ctx, D, args)
    nb = args.sf
    g = RandomArrayGenerator(ctx, 42)
    x = g.batches_dev(nb, JOIN_BATCH)
    pk = g.index_column_dev(nb, JOIN_BATCH)
    y = g.batches_dev(nb, JOIN_BATCH)
    fk = g.foreign_key_dev(JOIN_BATCH, nb, JOIN_BATCH)
    n = nb * JOIN_BATCH
    free, _ = torch.cuda.mem_get_info()
    full = ctx.join_ws_bytes(n, n)
    ws_bytes = min(full, max(free - (3 << 30), ctx.join_min_ws_bytes(n, n)))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
    out = torch.empty(3, dtype=torch.int64, device="cuda")
    res = {"workspace_gib": round(ws_bytes / 2**30, 2), "sliced": ws_bytes < full, "rows_per_side": n}
    # expected aggregates, computed with torch: every fk matches exactly one pk = its row number
    thr = 1 << 30
    flip = torch.tensor(-2**31, dtype=torch.int32, device="cuda")
    exp = {None: [0, 0, 0], thr: [0, 0, 0]}
    for s0 in range(0, n, 1 << 26):
        s1 = min(n, s0 + (1 << 26))
        yy = y[s0:s1].to(torch.int64) & 0xFFFFFFFF
        xx = x[(fk[s0:s1].to(torch.int64) & 0xFFFFFFFF)].to(torch.int64) & 0xFFFFFFFF
        keep = (y[s0:s1] ^ flip) < (thr - 2**31)
        exp[None][0] += s1 - s0
        exp[None][1] += int(yy.sum())
        exp[None][2] += int(xx.sum())
        exp[thr][0] += int(keep.sum())
        exp[thr][1] += int((yy * keep).sum())
        exp[thr][2] += int((xx * keep).sum())
    for name, t in (("join_sum", None), ("filter_join_sum", thr)):
        l0 = ctx.launches
        ms = timed_steps(D, lambda: ctx.join_aggr_dev(fk, y, pk, x, y_threshold=t, ws=ws, out=out),
                         args.steps, args.warmup)
        got = [int(v) for v in out.cpu().numpy().view("uint64")]
        if got != [v % (1 << 64) for v in exp[t]]:
            raise SystemExit(f"{name} self-check failed: {got} vs {exp[t]}")
        res[name] = {"ms_per_step": ms, "rows_per_s": n / (ms * 1e-3), "out_rows": got[0],
                     "launches_per_step": (ctx.launches - l0) // (args.steps + args.warmup)}
    del x, pk, y, fk, ws
    free_all()
    return res


def bench_join_aggr_sharded(ctx, D, args):
    """The fused pipeline over N GPUs: the join's exchange (P2PShuffleJoin: rows stored straight into the
    peers' receive buffers over NVLink), then every rank's local join adds the payloads of its output rows
    (b2_join_aggr_pairs_seg_cap_phased_dev) and ONE 24-byte all-reduce combines the ranks' sums. The predicate on
    L.y is applied in front of the link (b2_shuffle_p2p_count_lt_dev / _scatter_lt_dev): the rows that fail
    it are neither counted nor sent. Nothing is materialised."""
    import torch
    from dpu_olap_b200.generator import RandomArrayGenerator
    from dpu_olap_b200.sharded import P2PShuffleJoin
    G = D.world
    nb_total = args.sf
    if nb_total < G or nb_total * JOIN_BATCH > 1 << 32:
        return None
    first, nb = shard(nb_total, D)
    g = RandomArrayGenerator(ctx, 42)
    x = g.batches_dev(nb_total, JOIN_BATCH, take=(first, nb))
    pk = g.index_column_dev(nb_total, JOIN_BATCH, take=(first, nb))
    y = g.batches_dev(nb_total, JOIN_BATCH, take=(first, nb))
    fk = g.foreign_key_dev(JOIN_BATCH, nb_total, JOIN_BATCH, take=(first, nb))
    n = nb * JOIN_BATCH
    cap = n + n // 8 + 65536
    ok, pj, why = 1, None, None
    try:
        pj = P2PShuffleJoin(ctx, D.dist, D.rank, G, n, cap, n_build_total=nb_total * JOIN_BATCH)
    except Exception as e:  # noqa: BLE001 - agreed on by all ranks below
        ok, why = 0, f"{type(e).__name__}: {e}"[:200]
    if D.sum_int(ok) != G:
        return {"unavailable": why or "a peer rank could not set up symmetric memory"}
    jws_bytes = ctx.join_seg_cap_ws_bytes(cap, cap, pj.nr_expected, pj.skip, pj.seg_bits)
    jws = torch.empty(jws_bytes + 256, dtype=torch.uint8, device="cuda")
    part = torch.zeros(3, dtype=torch.int64, device="cuda")    # this rank's rows / sum_y / sum_x
    total = torch.zeros(3, dtype=torch.int64, device="cuda")   # all ranks' (int64 adds wrap mod 2^64 as uint64 sums do)
    thr = 1 << 30
    # expected aggregates with torch: fk matches exactly one pk = its global row number; pk batch b (and
    # with it x batch b) lives on rank b // nb, so every rank's x is broadcast once and gathered from
    flip = torch.tensor(-2**31, dtype=torch.int32, device="cuda")
    exp = {None: [0, 0, 0], thr: [0, 0, 0]}
    xbuf = torch.empty(n, dtype=torch.int32, device="cuda")
    chunk = 1 << 26
    for src in range(G):
        if src == D.rank:
            xbuf.copy_(x)
        D.dist.broadcast(xbuf, src=src)
        for s0 in range(0, n, chunk):
            s1 = min(n, s0 + chunk)
            f = (fk[s0:s1].to(torch.int64) & 0xFFFFFFFF) - src * n
            here = (f >= 0) & (f < n)
            keep = (y[s0:s1] ^ flip) < (thr - 2**31)
            xx = (xbuf[f.clamp(0, n - 1)].to(torch.int64) & 0xFFFFFFFF) * here
            exp[None][2] += int(xx.sum())
            exp[thr][2] += int((xx * keep).sum())
            del f, here, xx, keep
    for s0 in range(0, n, chunk):
        s1 = min(n, s0 + chunk)
        yy = y[s0:s1].to(torch.int64) & 0xFFFFFFFF
        keep = (y[s0:s1] ^ flip) < (thr - 2**31)
        exp[None][0] += s1 - s0
        exp[None][1] += int(yy.sum())
        exp[thr][0] += int(keep.sum())
        exp[thr][1] += int((yy * keep).sum())
        del yy, keep
    del xbuf

    def all_sum_u64(v: int) -> int:  # sums of up to 2^64 do not fit the int64 all-reduce: reduce the halves
        return ((D.sum_int(v >> 32) << 32) + D.sum_int(v & 0xFFFFFFFF)) % (1 << 64)

    res = {"rows_per_side": nb_total * JOIN_BATCH, "workspace_gib": round(jws_bytes / 2**30, 2), "sliced": False,
           "exchange": "P2PShuffleJoin (fused NVLink scatter, the predicate on L.y applied in front of the link), "
                       "local b2_join_aggr_pairs_seg_cap_phased_dev, one 24-byte all-reduce"}
    for name, t in (("join_sum", None), ("filter_join_sum", thr)):
        def local_join(l_buf, lseg, r_buf, rseg, nr_expected, seg_bits, skip_bits, abort, phase_bits, t=t):
            ctx.join_aggr_pairs_seg_cap_dev(l_buf, lseg, r_buf, rseg, nr_expected, seg_bits, skip_bits=skip_bits,
                                            ws=jws, out=part, y_threshold=t, abort=abort, phases=phase_bits)

        def step(t=t):
            pj.step(fk, y, pk, x, local_join, probe_lt=t)   # the predicate is applied in front of the link
            total.copy_(part)
            D.dist.all_reduce(total)
        l0 = ctx.launches
        ms = timed_steps(D, step, args.steps, args.warmup)
        got = [int(v) for v in total.cpu().numpy().view("uint64")]
        want = [all_sum_u64(v % (1 << 64)) for v in exp[t]]
        if got != want:
            raise SystemExit(f"sharded {name} self-check failed on rank {D.rank}: {got} vs {want}")
        res[name] = {"ms_per_step": ms, "rows_per_s": nb_total * JOIN_BATCH / (ms * 1e-3), "out_rows": got[0],
                     "launches_per_step": (ctx.launches - l0) // (args.steps + args.warmup), "self_check": "content"}
    del x, pk, y, fk, jws, pj
    free_all()
    return res


# ------------------------------------------------------------------------------------------------
# end-to-end (host buffers through the C ABI) and CPU baseline
# ------------------------------------------------------------------------------------------------
def bench_e2e_filter(ctx, D, args):
    """Same metric through the reference-facing host API: pinned host batches -> b2_filter_lt_u32_host
    -> b2_filter_fetch_host -> host result. H2D and D2H copies are inside the timed region."""
    import ctypes as C

    import numpy as np
    import torch

    from dpu_olap_b200._lib import Timings
    from dpu_olap_b200.generator import RandomArrayGenerator
    # every rank streams e2e_sf worth of batches through its own PCIe link (the host buffers of a
    # 64 GiB column do not fit this leg's time budget): total = e2e_sf x N
    e2e_sf = args.e2e_sf * D.world
    nb_total = e2e_sf << 7
    first, nb = shard(nb_total, D)
    n = nb * FILTER_BATCH
    g = RandomArrayGenerator(ctx, 42)
    dcol = g.batches_dev(nb_total, FILTER_BATCH, take=(first, nb))
    h_in = torch.empty(n, dtype=torch.int32, pin_memory=True)
    h_in.copy_(dcol)
    del dcol
    free_all()
    h_out = torch.empty(n, dtype=torch.int32, pin_memory=True)
    base_in, base_out = h_in.data_ptr(), h_out.data_ptr()
    ptrs = (C.c_void_p * nb)(*[base_in + 4 * FILTER_BATCH * b for b in range(nb)])
    lens = (C.c_int64 * nb)(*([FILTER_BATCH] * nb))
    counts = (C.c_int64 * nb)()
    total = C.c_uint64(0)
    t1 = Timings()
    lib, h = ctx._lib, ctx._h
    acc = {"h2d": 0, "d2h": 0, "launches": 0, "sel": 0}

    def step():
        ctx._ck(lib.b2_filter_lt_u32_host_into(h, ptrs, lens, nb, 1 << 30, base_out, n, counts,
                                               C.byref(total), C.byref(t1)), "b2_filter_lt_u32_host_into")
        acc["h2d"], acc["d2h"] = t1.h2d_bytes, t1.d2h_bytes
        acc["launches"], acc["sel"] = t1.kernel_launches, total.value

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(2, min(args.steps, 5))
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    ms = D.max_float((time.perf_counter() - t0) * 1e3) / steps
    # check the host result against numpy on the host copy of the input
    a = h_in.numpy().view(np.uint32)
    exp = a[a < (1 << 30)]
    got = h_out.numpy().view(np.uint32)[: acc["sel"]]
    if exp.size != got.size or not np.array_equal(exp, got):
        raise SystemExit("e2e filter self-check failed")
    rows = nb_total * FILTER_BATCH
    res = {"value": rows / (ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": acc["h2d"] * D.world,
           "d2h_bytes_per_step": acc["d2h"] * D.world, "ms_per_step": ms, "sf": e2e_sf, "rows": rows,
           "steps": steps, "scaling": "weak", "sf_per_gpu": args.e2e_sf,
           "workload": f"filter SF={args.e2e_sf} per GPU x {D.world} GPU(s) through host buffers (the 64 GiB SF=2048 "
                       "column does not fit a PCIe-bound leg's time budget); wall clock, max over ranks",
           "api": "b2_filter_lt_u32_host_into (pinned host batches in, pinned host result out; "
                  "upload / kernels / download of 64 MiB groups overlap)",
           "phases_ms": {"copy-to-dpu": t1.copy_to_dev_ms, "dpu-work": t1.dev_work_ms,
                         "copy-from-dpu": t1.copy_from_dev_ms},
           "gpu_launches_per_step": acc["launches"]}
    del h_in, h_out
    return res


def cpu_filter_sample(args, cpu_sf: int, seconds: float, steps: int | None = None, warmup: int = 1,
                      threads: int | None = None):
    """The reference's CPU path — FilterNative's Acero plan (filter_native.cc:36-84) on Arrow 24 via
    pyarrow — timed Prepare()+Run() per iteration as BM_Filter does (filter_benchmark.cc:30-49)."""
    import numpy as np
    import pyarrow as pa

    import oracle
    from oracle import arrow_native as an
    nb = cpu_sf << 7
    g = oracle.RandomArrayGenerator(42)
    seeds = [g.data_seed() for _ in range(nb)]
    flat = np.empty(nb * FILTER_BATCH, dtype=np.uint32)
    from concurrent.futures import ThreadPoolExecutor
    lib = oracle.oracle.lib()

    def fill(b):
        lib.orc_gen_u32(seeds[b], 0, 0xFFFFFFFF, FILTER_BATCH, flat.ctypes.data + 4 * FILTER_BATCH * b)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        list(ex.map(fill, range(nb)))
    batches = [flat[b * FILTER_BATCH:(b + 1) * FILTER_BATCH] for b in range(nb)]
    cores = threads or os.cpu_count() or 1
    pa.set_cpu_count(cores)
    times, sel = [], 0
    t_start = time.perf_counter()
    it = 0
    while True:
        f = an.FilterNative(batches)
        t0 = time.perf_counter()
        f.Prepare()
        sel = f.Run()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_start > seconds and len(times) >= 2:
            break
    rows = nb * FILTER_BATCH
    exp = int((flat < (1 << 30)).sum())
    if sel != exp:
        raise SystemExit("CPU filter self-check failed")
    ms = 1e3 * sum(times) / len(times)
    return {"value": rows / (ms * 1e-3), "unit": "rows/s", "cores": cores, "kind": "port",
            "sample": f"filter v<2^30 at SF={cpu_sf} ({nb} batches x 65536 rows = {rows * 4 / 2**30:.1f} GiB), "
                      f"{len(times)} timed runs of Prepare()+Run(), Arrow Acero {pa.__version__} via pyarrow "
                      f"(reference pins Arrow 8.0.0), {cores} threads",
            "ms_per_step": ms, "rows": rows, "selected": sel}



def _time_cpu(make_op, seconds: float, warmup: int = 1):
    """Prepare()+Run() per iteration, as the reference's BM_* loops time it; bounded by `seconds`."""
    times, res = [], None
    t_start = time.perf_counter()
    it = 0
    while True:
        op = make_op()
        t0 = time.perf_counter()
        op.Prepare()
        res = op.Run()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        it += 1
        if time.perf_counter() - t_start > seconds and len(times) >= 2:
            break
    return 1e3 * sum(times) / len(times), len(times), res


def _host_batches(t, batch: int):
    """A device column as the list of per-batch numpy views the *Native classes take."""
    import numpy as np
    a = t.cpu().numpy().view(np.uint32)
    return [a[i:i + batch] for i in range(0, a.size, batch)]


def cpu_ops_samples(ctx, args, seconds: float) -> dict:
    """the reference's other Native operators timed beside the GPU on this
    box's host cores — AggrNative (aggr_native.cc:39-93), TakeNative (take_native.cc:18-38),
    JoinNative (join_native.cc:14-102) — on a bounded sample (SF = --cpu-sf) of the same generator(42)
    workload, generated on the device (bit-identical) and copied to the host. Results are checked
    against the oracle / the GPU operator. Arrow Acero 24 via pyarrow, all host threads."""
    import numpy as np
    import pyarrow as pa

    import oracle
    from oracle import arrow_native as an
    from dpu_olap_b200.generator import RandomArrayGenerator
    sf = args.cpu_sf
    cores = os.cpu_count() or 1
    pa.set_cpu_count(cores)
    how = f"Arrow Acero {pa.__version__} via pyarrow (reference pins Arrow 8.0.0), {cores} threads, Prepare()+Run()"
    out = {}
    # ---- sum ----
    g = RandomArrayGenerator(ctx, 42)
    col = _host_batches(g.batches_dev(sf, SUM_BATCH), SUM_BATCH)
    ms, k, got = _time_cpu(lambda: an.AggrNative(col), seconds / 3)
    if got != sum(int(b.sum(dtype=np.uint64)) for b in col) % (1 << 64):
        raise SystemExit("CPU sum self-check failed")
    rows = sf * SUM_BATCH
    out["sum"] = {"value": rows / (ms * 1e-3), "unit": "rows/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                  "sample": f"AggrNative sum at SF={sf} ({sf} batches x {SUM_BATCH} rows), {k} timed runs, {how}"}
    del col
    # ---- take ----
    g = RandomArrayGenerator(ctx, 42)
    vals = _host_batches(g.batches_dev(sf, TAKE_BATCH), TAKE_BATCH)
    idx = _host_batches(g.batches_dev(sf, TAKE_IDX, 0, TAKE_BATCH - 1), TAKE_IDX)
    ms, k, got = _time_cpu(lambda: an.TakeNative(vals, idx), seconds / 3)
    if not np.array_equal(got[sf // 2], oracle.take(vals[sf // 2], idx[sf // 2])):
        raise SystemExit("CPU take self-check failed")
    nidx = sf * TAKE_IDX
    out["take"] = {"value": nidx / (ms * 1e-3), "unit": "indices/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                   "sample": f"TakeNative at SF={sf} ({sf} batches: {TAKE_BATCH} values, {TAKE_IDX} indices), "
                             f"{k} timed runs, {how}"}
    del vals, idx
    # ---- join ----
    g = RandomArrayGenerator(ctx, 42)
    xd = g.batches_dev(sf, JOIN_BATCH)
    pkd = g.index_column_dev(sf, JOIN_BATCH)
    yd = g.batches_dev(sf, JOIN_BATCH)
    fkd = g.foreign_key_dev(JOIN_BATCH, sf, JOIN_BATCH)
    exp = triple_checksum_torch(fkd, yd, xd[fkd.to(__import__("torch").int64) & 0xFFFFFFFF])
    x, pk, y, fk = (_host_batches(t, JOIN_BATCH) for t in (xd, pkd, yd, fkd))
    del xd, pkd, yd, fkd
    ms, k, tab = _time_cpu(lambda: an.JoinNative({"fk": fk, "y": y}, {"pk": pk, "x": x}), seconds, warmup=0)
    import torch
    cols = [torch.from_numpy(tab.column(c).combine_chunks().to_numpy().view(np.int32)) for c in ("fk", "y", "x")]
    if tab.num_rows != sf * JOIN_BATCH or triple_checksum_torch(*cols) != exp:
        raise SystemExit("CPU join self-check failed")
    rows = sf * JOIN_BATCH
    out["join"] = {"value": rows / (ms * 1e-3), "unit": "rows/s", "cores": cores, "kind": "port", "ms_per_step": ms,
                   "sample": f"JoinNative (Acero hashjoin inner fk = pk) at SF={sf}: {rows} rows per side, "
                             f"{k} timed runs, {how}"}
    return out


def _pinned(t):
    import torch
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t)
    return h


def bench_e2e_ops(ctx, D, args) -> dict:
    """e2e for sum / take / join: the same operators through the reference-facing host entry points
    (b2_sum_u32_host, b2_take_u32_host, b2_join_u32_host + b2_join_fetch_host — what SumGpu / TakeGpu /
    JoinGpu call) with PINNED HOST batches in and host results out, copies inside the timed region.
    One GPU, SF = --e2e-sf (bounded by PCIe time), wall clock over `steps` calls."""
    import ctypes as C

    import numpy as np
    import torch

    from dpu_olap_b200._lib import Timings
    from dpu_olap_b200.generator import RandomArrayGenerator
    sf = args.e2e_sf
    lib, h = ctx._lib, ctx._h
    steps = max(2, min(args.steps, 5))
    res = {}

    def table(*cols_and_batch):
        ptrs, lens = [], []
        for t, batch in cols_and_batch:
            for i in range(0, t.numel(), batch):
                ptrs.append(t.data_ptr() + 4 * i)
                lens.append(batch)
        return (C.c_void_p * len(ptrs))(*ptrs), (C.c_int64 * len(lens))(*lens)

    def timed(fn):
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / steps

    def leg(t: Timings, extra=None):
        d = {"h2d_bytes_per_step": int(t.h2d_bytes), "d2h_bytes_per_step": int(t.d2h_bytes),
             "phases_ms": {"copy-to-dpu": t.copy_to_dev_ms, "dpu-work": t.dev_work_ms,
                           "copy-from-dpu": t.copy_from_dev_ms},
             "gpu_launches_per_step": int(t.kernel_launches), "sf": sf, "steps": steps, "n_gpus": 1}
        if extra:
            for k in ("h2d_bytes_per_step", "d2h_bytes_per_step", "gpu_launches_per_step"):
                d[k] += {"h2d_bytes_per_step": int(extra.h2d_bytes), "d2h_bytes_per_step": int(extra.d2h_bytes),
                         "gpu_launches_per_step": int(extra.kernel_launches)}[k]
            for k, v in (("copy-to-dpu", extra.copy_to_dev_ms), ("dpu-work", extra.dev_work_ms),
                         ("copy-from-dpu", extra.copy_from_dev_ms)):
                d["phases_ms"][k] += v
        return d

    # ---- sum ----
    g = RandomArrayGenerator(ctx, 42)
    col = _pinned(g.batches_dev(sf, SUM_BATCH))
    ptrs, lens = table((col, SUM_BATCH))
    out, t1 = C.c_uint64(0), Timings()
    ms = timed(lambda: ctx._ck(lib.b2_sum_u32_host(h, ptrs, lens, sf, C.byref(out), C.byref(t1)), "b2_sum_u32_host"))
    if int(out.value) != int(col.numpy().view(np.uint32).sum(dtype=np.uint64)):
        raise SystemExit("e2e sum self-check failed")
    rows = sf * SUM_BATCH
    res["sum"] = {"value": rows / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms, "api": "b2_sum_u32_host", **leg(t1)}
    del col
    # ---- take ----
    g = RandomArrayGenerator(ctx, 42)
    vals = _pinned(g.batches_dev(sf, TAKE_BATCH))
    idx = _pinned(g.batches_dev(sf, TAKE_IDX, 0, TAKE_BATCH - 1))
    hout = torch.empty(sf * TAKE_IDX, dtype=torch.int32, pin_memory=True)
    vptrs, vlens = table((vals, TAKE_BATCH))
    iptrs, ilens = table((idx, TAKE_IDX))
    optrs, _ = table((hout, TAKE_IDX))
    ms = timed(lambda: ctx._ck(lib.b2_take_u32_host(h, vptrs, vlens, iptrs, ilens, sf, optrs, C.byref(t1)),
                               "b2_take_u32_host"))
    b = sf // 2
    ref = vals[b * TAKE_BATCH:(b + 1) * TAKE_BATCH][idx[b * TAKE_IDX:(b + 1) * TAKE_IDX].to(torch.int64)]
    if not torch.equal(ref, hout[b * TAKE_IDX:(b + 1) * TAKE_IDX]):
        raise SystemExit("e2e take self-check failed")
    nidx = sf * TAKE_IDX
    res["take"] = {"value": nidx / (ms * 1e-3), "unit": "indices/s", "ms_per_step": ms, "api": "b2_take_u32_host",
                   **leg(t1)}
    del vals, idx, hout
    # ---- join ----
    g = RandomArrayGenerator(ctx, 42)
    xd = g.batches_dev(sf, JOIN_BATCH)
    pkd = g.index_column_dev(sf, JOIN_BATCH)
    yd = g.batches_dev(sf, JOIN_BATCH)
    fkd = g.foreign_key_dev(JOIN_BATCH, sf, JOIN_BATCH)
    exp = triple_checksum_torch(fkd, yd, xd[fkd.to(torch.int64) & 0xFFFFFFFF])
    x, pk, y, fk = (_pinned(t) for t in (xd, pkd, yd, fkd))
    del xd, pkd, yd, fkd
    n = sf * JOIN_BATCH
    lptrs, llens = table((fk, JOIN_BATCH), (y, JOIN_BATCH))
    rptrs, rlens = table((pk, JOIN_BATCH), (x, JOIN_BATCH))
    o = [torch.empty(n, dtype=torch.int32, pin_memory=True) for _ in range(3)]
    nrows, t2 = C.c_uint64(0), Timings()

    def jstep():
        ctx._ck(lib.b2_join_u32_host(h, lptrs, llens, sf, rptrs, rlens, sf, C.byref(nrows), C.byref(t1)),
                "b2_join_u32_host")
        ctx._ck(lib.b2_join_fetch_host(h, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), n, C.byref(t2)),
                "b2_join_fetch_host")
    ms = timed(jstep)
    if int(nrows.value) != n or triple_checksum_torch(*[t.cuda() for t in o]) != exp:
        raise SystemExit("e2e join self-check failed")
    res["join"] = {"value": n / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms,
                   "api": "b2_join_u32_host + b2_join_fetch_host (JoinGpu::Run)", **leg(t1, t2),
                   "phases_note": "the build side's radix passes run under the probe side's upload (two "
                                  "streams), so dpu-work includes waiting for that upload and the phases "
                                  "do not add up to ms_per_step"}
    # the same join over DEVICE-RESIDENT columns (b2_col, what an Arrow consumer hands over as
    # ArrowDeviceArrays): the inputs are already in HBM, only the result columns cross PCIe
    from dpu_olap_b200.ops import DeviceColumn
    cols = [DeviceColumn.from_host(ctx, [t.numpy().view(np.uint32)]) for t in (fk, y, pk, x)]

    def dstep():
        outs = DeviceColumn.join(*cols)
        ptrs = [(C.c_void_p * 1)(t.data_ptr()) for t in o]
        for c, p in zip(outs, ptrs):
            ctx._ck(lib.b2_col_download_host(h, c._h, p, 1), "b2_col_download_host")
            c.close()
    ms_d = timed(dstep)
    if triple_checksum_torch(*[t.cuda() for t in o]) != exp:
        raise SystemExit("e2e device-column join self-check failed")
    res["join"]["device_inputs"] = {"value": n / (ms_d * 1e-3), "unit": "rows/s", "ms_per_step": ms_d,
                                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 12 * n,
                                    "api": "b2_join_u32_col over b2_col inputs (ArrowDeviceArray hand-over), "
                                           "b2_col_download_host for the three result columns"}
    for c in cols:
        c.close()
    del x, pk, y, fk, o
    free_all()
    return res


def bench_e2e_set_join(ctx, D, args) -> dict:
    """The sharded join through the C ABI's device set (b2_set_join_u32_host + b2_set_join_fetch_host:
    what the C++ JoinGpu calls with GpuSet::allocate(N)): ONE host process drives all N GPUs — upload
    of each GPU's batch range, count, device-side plan, NVLink scatter, local joins, download — with
    pinned host batches in and host columns out. Run by rank 0 alone while the other ranks wait."""
    import ctypes as C

    import torch

    from dpu_olap_b200._lib import Timings
    from dpu_olap_b200.generator import RandomArrayGenerator
    from dpu_olap_b200.ops import DeviceSet
    sf = args.e2e_sf * D.world
    g = RandomArrayGenerator(ctx, 42)
    xd = g.batches_dev(sf, JOIN_BATCH)
    pkd = g.index_column_dev(sf, JOIN_BATCH)
    yd = g.batches_dev(sf, JOIN_BATCH)
    fkd = g.foreign_key_dev(JOIN_BATCH, sf, JOIN_BATCH)
    exp = triple_checksum_torch(fkd, yd, xd[fkd.to(torch.int64) & 0xFFFFFFFF])
    x, pk, y, fk = (_pinned(t) for t in (xd, pkd, yd, fkd))
    del xd, pkd, yd, fkd
    free_all()
    n = sf * JOIN_BATCH

    def table(*cols):
        ptrs = [t.data_ptr() + 4 * i for t in cols for i in range(0, t.numel(), JOIN_BATCH)]
        return (C.c_void_p * len(ptrs))(*ptrs), (C.c_int64 * len(ptrs))(*([JOIN_BATCH] * len(ptrs)))
    lptrs, llens = table(fk, y)
    rptrs, rlens = table(pk, x)
    o = [torch.empty(n, dtype=torch.int32, pin_memory=True) for _ in range(3)]
    nrows, t1, t2 = C.c_uint64(0), Timings(), Timings()
    steps = max(2, min(args.steps, 5))
    with DeviceSet(list(range(D.world))) as ds:
        def step():
            ds._call("join_u32_host", lptrs, llens, sf, rptrs, rlens, sf, C.byref(nrows), C.byref(t1))
            ds._call("join_fetch_host", o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), n, C.byref(t2))
        for _ in range(2):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        ms = (time.perf_counter() - t0) * 1e3 / steps
    if int(nrows.value) != n or triple_checksum_torch(*[t.cuda() for t in o]) != exp:
        raise SystemExit("e2e set join self-check failed")
    res = {"value": n / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms, "sf": sf, "steps": steps, "n_gpus": D.world,
           "api": "b2_set_join_u32_host + b2_set_join_fetch_host (one process, N GPUs, CUDA events between devices)",
           "h2d_bytes_per_step": int(t1.h2d_bytes), "d2h_bytes_per_step": int(t1.d2h_bytes + t2.d2h_bytes),
           "phases_ms": {"copy-to-dpu": t1.copy_to_dev_ms, "dpu-work": t1.dev_work_ms,
                         "copy-from-dpu": t2.copy_from_dev_ms},
           "gpu_launches_per_step": int(t1.kernel_launches), "self_check": "content", "scaling": "weak"}
    del x, pk, y, fk, o
    return res


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    r = cpu_filter_sample(args, args.cpu_sf, 0.0, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": "filter rows/sec (v < 2^30, uint32 column)", "value": r["value"],
            "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic RandomArrayGenerator(42)",
            "config": {"workload": f"filter SF={args.sf} (each step = bounded sample at SF={args.cpu_sf})",
                       "batch_rows": FILTER_BATCH, "predicate": "v < 2^30"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def protect_stdout():
    """Libraries print banners on stdout (NCCL: "NCCL version ..."); the contract is ONE JSON line
    there. Everything else this process writes to fd 1 goes to stderr; emit() uses the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the operator path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    from dpu_olap_b200.ops import Context
    D = Dist(args.gpus)
    ctx = Context(D.local_rank)
    ops = [o for o in args.ops.split(",") if o]
    peak, peak_src = measured_peaks()

    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    thr25 = 1 << 30
    fres = bench_filter(ctx, D, args, [thr25])[thr25]
    clocks = sampler.stop() if D.rank == 0 else None

    extra = {}
    if "sweep" in ops:
        sweep = bench_filter(ctx, D, args, [42_949_673, 429_496_730, 1 << 31])
        extra["filter_selectivity_sweep"] = {
            f"{100 * v['selected'] / v['rows']:.0f}%": {
                "rows_per_s": v["rows_per_s"], "ms_per_step": v["ms_per_step"],
                "achieved_gbs": v["algorithmic_bytes"] / D.world / (v["ms_per_step"] * 1e-3) / 1e9}
            for v in sweep.values()}
    if "sum" in ops:
        extra["sum"] = bench_sum(ctx, D, args)
    if "take" in ops:
        extra["take"] = bench_take(ctx, D, args)
    if "join" in ops:
        extra["join"] = bench_join(ctx, D, args)
    if "nullable" in ops:
        extra["nullable"] = bench_nullable(ctx, D, args)
    if "joinsum" in ops:
        if D.world == 1:
            extra["join_aggregate"] = bench_join_aggr(ctx, D, args)
        else:
            # the newest multi-GPU path of the run: an allocation failure (the same on every rank: equal
            # shares) is reported in the line instead of ending it; a failed self-check still ends the run
            import torch
            try:
                extra["join_aggregate"] = bench_join_aggr(ctx, D, args)
            except torch.cuda.OutOfMemoryError as e:
                extra["join_aggregate"] = {"unavailable": f"out of memory: {e}"[:200]}
                free_all()
    if "wide" in ops:  # opt-in: the 64-bit aggregates and take
        extra["wide"] = bench_wide(ctx, D, args)
    e2e = None
    if not args.no_e2e:
        e2e = bench_e2e_filter(ctx, D, args)
        if D.world == 1:
            # the other operators through their host entry points (one GPU; the sharded host join is
            # b2_set_join_u32_host, timed by `--ops ...,setjoin`)
            for op, leg in bench_e2e_ops(ctx, D, args).items():
                if extra.get(op):
                    extra[op]["e2e"] = leg
        elif extra.get("join"):
            # N > 1: the sharded join behind the C ABI's device set, driven by rank 0's process alone
            # The other ranks wait on the CPU (c10d store), NOT in an NCCL barrier: a spinning collective
            # kernel of another process sharing a GPU with rank 0's kernels is exactly the co-residency
            # the B200 profiling notes warn about (context-switch timeouts).
            D.barrier()
            from datetime import timedelta
            store = D.dist.distributed_c10d._get_default_store()
            if D.rank == 0:
                try:
                    extra["join"]["e2e"] = bench_e2e_set_join(ctx, D, args)
                except Exception as e:  # noqa: BLE001 - the device-resident numbers above stand on their own
                    extra["join"]["e2e"] = {"value": None, "error": f"{type(e).__name__}: {e}"[:300]}
                store.set("b2_set_join_done", "1")
            else:
                store.wait(["b2_set_join_done"], timedelta(minutes=20))
            D.barrier()
    cpu = None
    if D.rank == 0 and D.world == 1 and not args.no_cpu:
        cpu = cpu_filter_sample(args, args.cpu_sf, args.cpu_seconds)
        # the reference's own default is half the hardware threads (host/system.h:20): report that too
        half = max(1, (os.cpu_count() or 2) // 2)
        h = cpu_filter_sample(args, args.cpu_sf, args.cpu_seconds / 3, threads=half)
        cpu["half_cores"] = {"value": h["value"], "unit": h["unit"], "cores": half, "ms_per_step": h["ms_per_step"]}
        for op, base in cpu_ops_samples(ctx, args, args.cpu_seconds).items():
            if extra.get(op):
                extra[op]["cpu_baseline"] = base
    # per-operator rooflines (against the measured copy bandwidth and the nominal 8 TB/s)
    if extra.get("sum"):
        r = extra["sum"]
        r["roofline"] = roofline_block("sum", 4.0 * r["rows"], r["ms_per_step"], peak, peak_src, r["rows"], D.world,
                                       "4 B read per row", "sum_u32_kernel")
    if extra.get("take"):
        r = extra["take"]
        r["roofline"] = roofline_block("take", 12.0 * r["indices"], r["ms_per_step"], peak, peak_src, r["indices"],
                                       D.world, "12 B per index (index + value + output); the DRAM traffic is the "
                                       "16 MiB batch window, see window_gbs", "take_u32_vec_kernel")
    if extra.get("join"):
        r = extra["join"]
        r["roofline"] = roofline_block(
            "join", float(r["algorithmic_bytes_per_row"]) * r["rows_per_side"], r["ms_per_step"], peak, peak_src,
            r["rows_per_side"], D.world,
            "single-pass ideal 8 B per build row + 8 B per probe row + 12 B per output row; the two-pass radix "
            "design is modelled at 116 B per row pair (DESIGN.md section 3); time = whole join step "
            "(all radix passes + probe" + (", + NVLink shuffle" if D.world > 1 else "") + ")",
            "part_scatter_*_kernel x4 + part_hist_kernel x4 + join_probe_kernel")
        jt = ncu_traffic("join")
        if jt and jt.get("kernels") and D.world == 1:
            # per-kernel times and DRAM rates of the same command under ncu (cold-cache, serialised launches:
            # the SHARE of the step is what carries over, profiles/r2_launches_bench_sf2048.csv)
            r["roofline"]["kernels_ncu"] = jt["kernels"]
            r["roofline"]["kernels_ncu_source"] = jt.get("source")
        n1 = ncu_traffic("join_n1_ms")
        if n1 and D.world > 1 and n1.get(str(args.sf)):
            r["speedup_vs_n1"] = n1[str(args.sf)] / r["ms_per_step"]
            r["n1_ms_source"] = n1.get("source")

    if D.rank == 0:
        achieved = fres["algorithmic_bytes"] / D.world / (fres["ms_per_step"] * 1e-3) / 1e9
        line = {
            "metric": "filter rows/sec (v < 2^30, uint32 column)",
            "value": fres["rows_per_s"], "unit": "rows/s", "n_gpus": D.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": fres["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic: RandomArrayGenerator(42) generated on the device (bit-identical to the reference's host generator)",
            "config": {"workload": f"filter SF={args.sf}: {args.sf << 7} batches x {FILTER_BATCH} uint32 rows "
                                   f"({fres['rows'] * 4 / 2**30:.0f} GiB), predicate v < 2^30 "
                                   f"({100 * fres['selected'] / fres['rows']:.1f} % selected)",
                       "sf": args.sf, "rows": fres["rows"], "rows_per_gpu": fres["rows_per_rank"],
                       "sharding": "contiguous batch ranges per GPU, no collective",
                       "l2": "inputs (>= 8 GiB per GPU) far exceed the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         # DRAM bytes per launch per GPU from the ncu capture committed for this round
                         # (profiles/r2_traffic.json names the CSV), scaled to this launch's rows
                         **filter_traffic(fres, D.world),
                         "algorithmic_bytes_per_launch": fres["algorithmic_bytes"] / D.world,
                         "peak_source": peak_src,
                         "kernel": "filter_lt_u32_kernel",
                         "algorithmic_bytes_per_row": fres["algorithmic_bytes"] / fres["rows"],
                         "note": "bytes = 4 B read per row + 4 B written per selected row, per GPU; "
                                 "time = whole step (descriptor memset + filter kernel + batch-end kernel)"},
            "gpu_launches": fres["launches_per_step"] * args.steps,
            "clocks": clocks,
            "e2e": e2e if e2e else {"value": None, "unit": "rows/s", "h2d_bytes_per_step": 0,
                                    "d2h_bytes_per_step": 0, "skipped": True},
            "cpu_baseline": cpu,
            "ops": extra,
        }
        emit(line)
    ctx.close()
    D.close()


if __name__ == "__main__":
    main()
