#!/usr/bin/env python
"""Print the headline fields of bench.py JSON lines:  python tools/show_bench.py gpurun_out/*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e)
        continue
    r = j.get("roofline") or {}
    print("%-40s N=%s value=%.3fM ms=%.4f e2e=%.3fM frac=%s g1_ms=%s loss=%s" % (
        f.split("/")[-1], j.get("n_gpus"), j["value"] / 1e6, j["ms_per_step"], (j.get("e2e") or {}).get("value", 0) / 1e6,
        ("%.3f" % r["frac"]) if r.get("frac") else None, r.get("avg_launch_ms"), j.get("loss")))
    if j.get("kernels_ms") and "-k" in sys.argv:
        print("   ", {k: round(v, 4) for k, v in j["kernels_ms"].items()})
