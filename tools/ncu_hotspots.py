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
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
        lines, fname, hdr2 = [], "", None
        for r in csv.reader(out.splitlines()):
            if not r:
                continue
            if r[0] == "File Path":
                fname = r[1].split("/")[-1]
            elif r[0] == "Line No":
                hdr2 = r
            elif hdr2 and r[0].isdigit():
                ns, ne = hdr2.index("# Samples"), hdr2.index("Instructions Executed")
                try:
                    lines.append((int(r[ns] or 0), int(r[ne] or 0), fname, int(r[0]), r[1].strip()))
                except ValueError:
                    pass
        print("hottest source lines (file:line, samples, share, warp instructions executed, text):")
        for sm, ne, fn, ln, txt in sorted(sorted(lines, reverse=True)[:top], key=lambda x: (x[2], x[3])):
            print("  %-18s %6d %5.1f%% %10d  %s" % ("%s:%d" % (fn, ln), sm, 100.0 * sm / max(total, 1), ne, txt[:100]))
    except Exception as e:  # noqa: BLE001
        print("(no CUDA source view: %s)" % e)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
