#!/usr/bin/env python3
"""Per-source-line table of an .ncu-rep taken with --import-source on: every line of one file with warp-stall samples,
warp instructions executed and the dominant stall reason, optionally summed over line ranges.
  python tools/ncu_lines.py file.ncu-rep source.cu [lo-hi:label ...]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, fname, ranges):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr = "", None
    lines = defaultdict(lambda: [0, 0, defaultdict(int), ""])
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and cur == fname:
            ns, ne = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:
                e = lines[int(r[0])]
                e[0] += int(r[ns] or 0)
                e[1] += int(r[ne] or 0)
                e[3] = r[1].strip()
                for i, h in enumerate(hdr):
                    if h.startswith("stall_") and r[i]:
                        try:
                            e[2][h[6:]] += int(r[i])
                        except ValueError:
                            pass
            except ValueError:
                pass
    ts = sum(e[0] for e in lines.values()) or 1
    ti = sum(e[1] for e in lines.values()) or 1
    print("%s: %d samples, %d warp instructions in %s" % (path, ts, ti, fname))
    if ranges:
        for spec in ranges:
            rng, label = spec.split(":")
            lo, hi = (int(x) for x in rng.split("-"))
            s = sum(e[0] for ln, e in lines.items() if lo <= ln <= hi)
            i = sum(e[1] for ln, e in lines.items() if lo <= ln <= hi)
            reasons = defaultdict(int)
            for ln, e in lines.items():
                if lo <= ln <= hi:
                    for k, v in e[2].items():
                        reasons[k] += v
            top = ", ".join("%s %d" % kv for kv in sorted(reasons.items(), key=lambda x: -x[1])[:4])
            print("  %-28s lines %5d-%5d  samples %6d %5.1f%%   inst %10d %5.1f%%   %s" % (label, lo, hi, s, 100.0 * s / ts, i, 100.0 * i / ti, top))
    else:
        for ln in sorted(lines):
            e = lines[ln]
            if e[0] or e[1]:
                dom = max(e[2].items(), key=lambda x: x[1])[0] if e[2] else ""
                print("  %5d %6d %5.1f%% %10d %5.1f%%  %-16s %s" % (ln, e[0], 100.0 * e[0] / ts, e[1], 100.0 * e[1] / ti, dom, e[3][:110]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
