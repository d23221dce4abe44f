#!/usr/bin/env python3
"""Brief of an .ncu-rep: per kernel name (the longest launch of each) duration, DRAM bytes, throughput, occupancy, issue,
top stalls.   python tools/ncu_brief.py file.ncu-rep [...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print(path, ": no launches")
        return
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    kn, dur = ci.get("Kernel Name"), ci.get("gpu__time_duration.sum")
    best, count = {}, {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        n = short(r[kn])
        count[n] = count.get(n, 0) + 1
        try:
            d = float(r[dur].replace(",", ""))
        except ValueError:
            d = 0.0
        if n not in best or d > best[n][0]:
            best[n] = (d, r)
    print(path)
    for n in sorted(best):
        r = best[n][1]
        print("%s   (%d launch%s captured; the longest)" % (n, count[n], "" if count[n] == 1 else "es"))
        for k in KEYS:
            if k in ci:
                print("  %-62s %s %s" % (k, r[ci[k]], units[ci[k]]))
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h:
                try:
                    st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  stalls (warps per issue):", ", ".join("%s %.2f" % (nm, v) for v, nm in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
