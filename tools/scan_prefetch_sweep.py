#!/usr/bin/env python
"""Exclusive scan of 2^28 uint32: time per scan variant and L2 prefetch distance (param scan_prefetch_tiles)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda.radixsort_b200 as rs
ap = argparse.ArgumentParser()
ap.add_argument("--variants", default="8")
ap.add_argument("--distances", default="0,128,256,512,1024,2048")
ap.add_argument("--log2n", type=int, default=28)
args = ap.parse_args()
rs.load()
n = 1 << args.log2n
x = (torch.arange(n, dtype=torch.int64, device="cuda") % 3).to(torch.int32)
out = torch.empty_like(x)
ws = rs.Workspace("cuda")
for v in [int(t) for t in args.variants.split(",")]:
    rs.set_param("scan_variant", v)
    for d in [int(t) for t in args.distances.split(",")]:
        rs.set_param("scan_prefetch_tiles", d)
        for _ in range(3):
            rs.exclusive_scan(x, out=out, workspace=ws)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            rs.exclusive_scan(x, out=out, workspace=ws)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(json.dumps({"scan_variant": v, "prefetch_tiles": d, "ms": round(ms, 4), "gbs": round(8 * n / ms / 1e6, 1)}), flush=True)
ref = torch.cumsum(x.to(torch.int64), 0) - x
print(json.dumps({"ok": bool(torch.equal(out.to(torch.int64) & 0xFFFFFFFF, ref & 0xFFFFFFFF))}))
