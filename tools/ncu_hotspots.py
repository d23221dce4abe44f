#!/usr/bin/env python3
"""Source-level view of an .ncu-rep (first kernel): hottest SASS instructions by warp-stall samples with their dominant
stall reason and execution counts, plus per-source-line totals when the report was taken with --import-source on.
  python tools/ncu_hotspots.py file.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def page(path, view):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # first row: kernel name; second: header
    start = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#", "Line"))
    return rows[start - 1][1] if start else "", rows[start], rows[start + 1:]


def main(path, top=40):
    name, hdr, rows = page(path, "sass")
    ci = {h: i for i, h in enumerate(hdr)}
    samp, execd, src = ci["# Samples"], ci["Instructions Executed"], ci["Source"]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    total = sum(int(r[samp] or 0) for r in rows)
    tot_inst = sum(int(r[execd] or 0) for r in rows)
    print(path)
    print(name[:120])
    print("warp-stall samples %d, warp instructions executed %d, SASS instructions %d" % (total, tot_inst, len(rows)))
    by_reason = defaultdict(int)
    for r in rows:
        for i in stall_cols:
            try:
                by_reason[hdr[i]] += int(r[i] or 0)
            except ValueError:
                pass
    if by_reason:
        s = sum(by_reason.values()) or 1
        print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in sorted(by_reason.items(), key=lambda x: -x[1])[:8]))
    print("hottest SASS instructions (index, samples, share, executions, dominant stall, instruction):")
    order = sorted(range(len(rows)), key=lambda k: -int(rows[k][samp] or 0))[:top]
    for k in sorted(order):
        r = rows[k]
        dom = ""
        if stall_cols:
            best = max(stall_cols, key=lambda i: int(r[i] or 0))
            dom = hdr[best][6:]
        print("  %5d %6d %5.1f%% %8s  %-16s %s" % (k, int(r[samp] or 0), 100.0 * int(r[samp] or 0) / max(total, 1), r[execd], dom, r[src].strip()))
    try:
        _, hdr2, rows2 = page(path, "cuda")
        c2 = {h: i for i, h in enumerate(hdr2)}
        if "# Samples" in c2:
            print("hottest source lines (line, samples, share, instructions executed, text):")
            ls = sorted(rows2, key=lambda r: -int(r[c2["# Samples"]] or 0))[:top]
            for r in sorted(ls, key=lambda r: int(r[0]) if r[0].isdigit() else 0):
                print("  %6s %6s %5.1f%% %9s  %s" % (r[0], r[c2["# Samples"]], 100.0 * int(r[c2["# Samples"]] or 0) / max(total, 1),
                                                 r[c2["Instructions Executed"]], r[c2["Source"]].strip()[:110]))
    except Exception as e:  # noqa: BLE001
        print("(no CUDA source view: %s)" % e)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
