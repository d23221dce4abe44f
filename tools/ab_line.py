#!/usr/bin/env python3
"""One summary line of a bench.py JSON (A/B runs of tools/gpu_round.sh)."""
import json
import sys

try:
    j = json.load(open(sys.argv[1]))
    r = j["roofline"]
    print("AB %s  value %.0f  chain %s  iso %s" % (
        sys.argv[2], j["value"], {k: round(v, 1) for k, v in r["chain_us_per_frame"].items()},
        {k: round(v, 1) for k, v in r["isolated"]["chain_us_per_frame"].items()}))
except Exception as e:  # noqa: BLE001
    print("AB", sys.argv[2], "failed:", e)
