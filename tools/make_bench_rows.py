"""Rows of bench.py's parity check on one GPU for a grid no CPU here can hold (1024x1024x512): writes
gpurun_out/bench_rows_<wl>.json; merge it into tests/golden/bench_rows.json.     python tools/make_bench_rows.py c5"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dotsocp_b200 as dp
from dotsocp_b200 import slab
wl = sys.argv[1] if len(sys.argv) > 1 else "c5"
out = bench.parity_rows(dp, slab, wl, 0, 1, lambda: None)
rows = out.get("kkt")
if rows is None:
    print("golden already present:", out)
else:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"bench_rows_{wl}.json"), "w") as f:
        json.dump({wl: {"source": "single-GPU run of this library (tools/make_bench_rows.py); cross-checked against the CPU oracle at 512x512x256",
                        "kkt": rows, "priVal": out["priVal"]}}, f)
    print("rows written", len(rows), rows[-1])
