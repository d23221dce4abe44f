#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of the built objects the memory-instruction mix and the lines that prove the TMA
paths (UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier).
    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"orb.o": ["k_fast_cells"], "surfel.o": ["k_fuse_pipeILi3ELb1", "k_fuse_streamILi3ELb1", "k_fuse_oneILi4ELi1ELb1ELb0", "k_sp_fit2", "k_sp_seeds2"]}
MEM = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOMG", "ATOMS", "RED", "UTMALDG", "UBLKCP", "SYNCS", "CCTL", "LDC", "MUFU", "BAR", "SHFL", "VOTE")


def main():
    for obj, kernels in WANT.items():
        path = os.path.join(ROOT, "manhattanslam_b200", "build", obj)
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        blocks = re.split(r"\n\s*Function : ", out)
        for k in kernels:
            for b in blocks[1:]:
                name = b.split("\n", 1)[0]
                if k not in name:
                    continue
                ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(.*?);", b)
                hist = collections.Counter()
                for i in ins:
                    op = i.split()[1] if i.startswith("@") else i.split()[0]
                    base = op.split(".")[0]
                    if base in MEM:
                        hist[op if base in ("LDG", "STG", "LDS", "STS", "UTMALDG", "UBLKCP", "SYNCS") else base] += 1
                dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip() or name
                dem = dem.replace("(anonymous namespace)::", "").split("(")[0]
                print("== %s  (%s, %d SASS instructions)" % (dem, obj, len(ins)))
                print("   memory / sync instruction mix:", ", ".join("%s x%d" % kv for kv in sorted(hist.items())))
                seen = collections.OrderedDict()
                for i in ins:
                    if re.search(r"UTMALDG|UBLKCP|SYNCS\.(ARRIVE|PHASECHK|EXCH)", i):
                        seen[i.strip()] = seen.get(i.strip(), 0) + 1
                for line, c in seen.items():
                    print("     ", line + ("   (x%d)" % c if c > 1 else ""))
                break


if __name__ == "__main__":
    main()
