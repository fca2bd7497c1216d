#!/usr/bin/env python
"""Time the whole sort for several L2 prefetch distances of the column-sweep kernel (param prefetch_tiles).
    python tools/prefetch_sweep.py [--variant 95] [--distances 0,148,296,444,592]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda.radixsort_b200 as rs

ap = argparse.ArgumentParser()
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--distances", default="0,74,148,222,296,444,592,888")
ap.add_argument("--log2n", type=int, default=28)
ap.add_argument("--workload", default="uniform")
args = ap.parse_args()
rs.load()
n = 1 << args.log2n
pairs = args.workload == "pairs"
keys = rs.generate("uniform" if pairs else args.workload, n)
vals = torch.arange(n, dtype=torch.int32, device="cuda") if pairs else None
vout = torch.empty_like(keys) if pairs else None
def sort():
    if pairs:
        rs.sort_pairs(keys, vals, 8, out_keys=out, out_vals=vout, workspace=ws)
    else:
        rs.sort_keys(keys, 8, out=out, workspace=ws)
out = torch.empty_like(keys)
ws = rs.Workspace("cuda")
rs.set_param("variant", args.variant)
for d in [int(x) for x in args.distances.split(",")]:
    rs.set_param("prefetch_tiles", d)
    for _ in range(3):
        sort()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        sort()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(json.dumps({"variant": rs.get_param("effective_variant"), "prefetch_tiles": d, "ms": round(ms, 4),
                      "gkeys_s": round(n / ms / 1e6, 2)}), flush=True)
