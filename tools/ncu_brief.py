#!/usr/bin/env python3
"""Brief of an .ncu-rep (first launch): duration, DRAM bytes, throughput, occupancy, issue, top stalls."""
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    g = lambda k: (r[hdr.index(k)] + " " + units[hdr.index(k)]) if k in hdr else "n/a"
    print(path)
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
              "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
              "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]:
        print("  %-62s %s" % (k, g(k)))
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stalls (warps per issue):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
