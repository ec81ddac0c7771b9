#!/usr/bin/env python
"""launch_summary.py — one line per kernel launch of an `ncu --csv --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active...` log.

    python tools/launch_summary.py profiles/<launches>.csv [--min-ms 0.05] [--last-step N]
"""
import argparse, collections, csv


def main():
    p = argparse.ArgumentParser()
    p.add_argument("csv")
    p.add_argument("--min-ms", type=float, default=0.05)
    p.add_argument("--tail", type=int, default=0, help="only the last N launches")
    a = p.parse_args()
    rows = list(csv.reader(l for l in open(a.csv) if l.startswith('"')))
    hdr, d = rows[0], collections.OrderedDict()
    for r in rows[1:]:
        rec = dict(zip(hdr, r))
        e = d.setdefault(int(rec["ID"]), {"k": rec["Kernel Name"], "grid": rec["Grid Size"], "blk": rec["Block Size"]})
        e[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    items = list(d.items())
    if a.tail:
        items = items[-a.tail:]
    tot_ms = tot_b = 0.0
    for i, v in items:
        ms = v.get("gpu__time_duration.sum", 0) / 1e6
        rd, wr = v.get("dram__bytes_read.sum", 0) / 1e9, v.get("dram__bytes_write.sum", 0) / 1e9
        tot_ms += ms
        tot_b += rd + wr
        if ms >= a.min_ms:
            name = v["k"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            print(f"{i:4d} {ms:8.3f} ms  rd {rd:6.2f} wr {wr:6.2f} GB  {(rd + wr) / ms:6.2f} TB/s  "
                  f"inst {v.get('smsp__inst_executed.sum', 0) / 1e9:6.3f} G  issue "
                  f"{v.get('smsp__issue_active.avg.pct_of_peak_sustained_elapsed', 0):5.1f} %  {name} {v['grid']}x{v['blk']}")
    print(f"total {tot_ms:.3f} ms, {tot_b:.1f} GB DRAM")


if __name__ == "__main__":
    main()
