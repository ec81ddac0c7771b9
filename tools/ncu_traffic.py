#!/usr/bin/env python
"""ncu_traffic.py — turn ncu launch-list CSVs into profiles/r2_traffic.json, the file bench.py reads
for `roofline.traffic` (DRAM bytes actually moved, never a constant typed into bench.py).

    python tools/ncu_traffic.py --filter profiles/<csv> --filter-rows 2**24 --filter-selected N \
                                --join profiles/<csv> --join-rows 2**31 [--sum ... --take ...]

Each CSV is the `--csv --log-file` output of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ...
The operator's kernels are selected by name; bytes of all matching launches are added up and divided
by (steps x rows) -> DRAM bytes per row. `--steps` = operator invocations inside the capture.
"""
from __future__ import annotations

import argparse
import csv
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

KERNELS = {
    "filter": r"filter_lt_u32_kernel|filter_batch_end|filter_finish",
    "sum": r"sum_u32_kernel",
    "take": r"take_u32",
    "join": r"part_|join_|exclusive_scan|scan_",
}


def read_launches(path: Path):
    """-> list of {kernel, metric: value} per launch id, from ncu's long-format CSV."""
    rows = {}
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(r["ID"], {"kernel": r["Kernel Name"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9,
                 "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        d[r["Metric Name"]] = v * scale
    return list(rows.values())


def summarise(path: Path, pattern: str, rows: float, steps: int) -> dict:
    sel = [l for l in read_launches(path) if re.search(pattern, l["kernel"])]
    rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in sel)
    wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in sel)
    ns = sum(l.get("gpu__time_duration.sum", 0.0) for l in sel)
    return {"launches": len(sel), "steps": steps, "rows_per_step": rows, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_bytes_per_row": (rd + wr) / (rows * steps), "kernel_ms_per_step_under_ncu": ns / 1e6 / steps,
            "source": str(path.relative_to(ROOT)) if path.is_absolute() else str(path)}


def per_kernel(path: Path, pattern: str, steps: int) -> dict:
    """name -> launches per step, mean ms per launch, DRAM GB per launch, TB/s of DRAM traffic (kernels of
    at least 0.05 ms), for the operator's per-kernel table in the bench line."""
    agg: dict = {}
    for l in read_launches(path):
        if not re.search(pattern, l["kernel"]):
            continue
        name = re.sub(r"\(.*", "", l["kernel"]).replace("<unnamed>::", "").replace("void ", "").strip()
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += l.get("gpu__time_duration.sum", 0.0)
        a[2] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    out = {}
    for name, (n, ns, by) in agg.items():
        if ns / n < 5e4:
            continue
        out[name] = {"launches_per_step": n / steps, "ms_per_launch": round(ns / n / 1e6, 3),
                     "dram_gb_per_launch": round(by / n / 1e9, 2), "dram_tbs": round(by / ns / 1e3, 2)}
    return out


def main():
    p = argparse.ArgumentParser()
    for op in KERNELS:
        p.add_argument(f"--{op}")
        p.add_argument(f"--{op}-rows", default="0")
        p.add_argument(f"--{op}-steps", type=int, default=1)
    p.add_argument("--filter-selected", default="0", help="selected rows per step (algorithmic bytes = 4 rows + 4 selected)")
    p.add_argument("--join-n1-ms", default=None, help="sf:ms[,sf:ms] of this round's one-GPU join, for speedup_vs_n1")
    p.add_argument("--join-n1-source", default=None)
    p.add_argument("--out", default=str(ROOT / "profiles" / "r2_traffic.json"))
    a = p.parse_args()
    out_path = Path(a.out)
    out = json.loads(out_path.read_text()) if out_path.exists() else {}
    for op, pat in KERNELS.items():
        f = getattr(a, op)
        if not f:
            continue
        rows = float(eval(getattr(a, f"{op}_rows"), {}))  # noqa: S307 - "2**31" on my own command line
        s = summarise(Path(f), pat, rows, getattr(a, f"{op}_steps"))
        if op == "filter":
            sel = float(eval(a.filter_selected, {}))  # noqa: S307
            alg = 4 * rows + 4 * sel
            s["dram_bytes_per_algorithmic_byte"] = (s["dram_read_bytes"] + s["dram_write_bytes"]) / (alg * s["steps"])
        if op == "join":
            s["kernels"] = per_kernel(Path(f), pat, getattr(a, f"{op}_steps"))
        out[op] = s
    if a.join_n1_ms:
        out["join_n1_ms"] = {k: float(v) for k, v in (kv.split(":") for kv in a.join_n1_ms.split(","))}
        out["join_n1_ms"]["source"] = a.join_n1_source
    try:
        out["git"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True,
                                    text=True).stdout.strip()
    except Exception:
        pass
    out_path.write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
