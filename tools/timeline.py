#!/usr/bin/env python3
"""Reads the chrome trace bench.py writes with MSL_TIMELINE=<file> (CUPTI activity records of a few steps) and prints, per
step, when each stream's kernels ran: the fuse chain's span, the superpixel stage's span, how long the chain of batch k+1
waited after the chain of batch k, and the time during which only low-occupancy kernels were on the device."""
import json
import sys
from collections import defaultdict


def short(name):
    n = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if n.startswith("void "):
        n = n[5:]
    return n.split("<")[0].split("(")[0].strip() or name[:24]


def main(path):
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    by_stream = defaultdict(list)
    for e in ev:
        by_stream[e["args"].get("stream")].append(e)
    print("streams:")
    for st, es in sorted(by_stream.items(), key=lambda kv: -len(kv[1])):
        names = defaultdict(int)
        for e in es:
            names[short(e["name"])[:24]] += 1
        busy = sum(e["dur"] for e in es)
        print("  stream %s: %d events, busy %.2f ms, span %.2f..%.2f ms: %s" % (st, len(es), busy / 1000, (es[0]["ts"] - t0) / 1000,
              (es[-1]["ts"] + es[-1]["dur"] - t0) / 1000, ", ".join("%s x%d" % kv for kv in sorted(names.items(), key=lambda kv: -kv[1])[:6])))
    fuse = [e for e in ev if "k_fuse_pipe" in e["name"]]
    if not fuse:
        return
    # chains: runs of 64 consecutive fuse launches
    gaps = [(fuse[i + 1]["ts"] - (fuse[i]["ts"] + fuse[i]["dur"])) for i in range(len(fuse) - 1)]
    durs = [e["dur"] for e in fuse]
    print("k_fuse_pipe: %d launches, avg dur %.1f us, median gap %.1f us" % (len(fuse), sum(durs) / len(durs), sorted(gaps)[len(gaps) // 2]))
    big = [(i, g) for i, g in enumerate(gaps) if g > 100]
    for i, g in big:
        a, b = fuse[i]["ts"] + fuse[i]["dur"], fuse[i + 1]["ts"]
        inside = defaultdict(float)
        for e in ev:
            s, t = max(e["ts"], a), min(e["ts"] + e["dur"], b)
            if t > s:
                inside[short(e["name"])[:20]] += t - s
        print("  gap after launch %d: %.0f us (at %.2f ms); running inside: %s" % (i, g, (a - t0) / 1000,
              ", ".join("%s %.0f" % kv for kv in sorted(inside.items(), key=lambda kv: -kv[1])[:8])))
    # per-launch duration along a chain (does the kernel slow down while other streams are busy?)
    n = len(fuse)
    for c0 in range(0, n, 64):
        ch = fuse[c0:c0 + 64]
        print("  chain %d: start %.2f ms, end %.2f ms, span %.2f ms, sum dur %.2f ms, first8 avg %.1f us, last8 avg %.1f us" % (
            c0 // 64, (ch[0]["ts"] - t0) / 1000, (ch[-1]["ts"] + ch[-1]["dur"] - t0) / 1000, (ch[-1]["ts"] + ch[-1]["dur"] - ch[0]["ts"]) / 1000,
            sum(e["dur"] for e in ch) / 1000, sum(e["dur"] for e in ch[:8]) / 8, sum(e["dur"] for e in ch[-8:]) / 8))
    sp = [e for e in ev if "k_sp_" in e["name"]]
    # superpixel stages: split where the gap between consecutive k_sp_init launches lies
    inits = [e for e in sp if "k_sp_init" in e["name"]]
    recs = [e for e in sp if "k_sp_records" in e["name"]]
    for a, b in zip(inits, recs):
        print("  superpixel stage: k_sp_init at %.2f ms .. k_sp_records end %.2f ms (%.2f ms)" % ((a["ts"] - t0) / 1000, (b["ts"] + b["dur"] - t0) / 1000,
              (b["ts"] + b["dur"] - a["ts"]) / 1000))


if __name__ == "__main__":
    main(sys.argv[1])
