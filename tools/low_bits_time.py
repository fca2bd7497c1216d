#!/usr/bin/env python
"""Time b200sort_keys_low_bits against the full sort on keys that share their top bits (one shard of the multi-GPU sort)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda.radixsort_b200 as rs
rs.load()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 29
n = 1 << log2n
keys = rs.generate("uniform", n)
out = torch.empty_like(keys)
ws = rs.Workspace("cuda")
for kb in (32, 31, 30, 29, 28, 25, 24):
    k = (keys & ((1 << kb) - 1 if kb < 32 else -1)) if kb < 32 else keys
    if kb < 32:
        k = (k.to(torch.int64) & ((1 << kb) - 1)).to(torch.int32)
    for _ in range(2):
        rs.sort_keys(k, 8, out=out, workspace=ws, key_bits=kb)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        rs.sort_keys(k, 8, out=out, workspace=ws, key_bits=kb)
    b.record()
    torch.cuda.synchronize()
    bad = rs.verify(out)[0]
    print(json.dumps({"n": n, "key_bits": kb, "ms": round(a.elapsed_time(b) / 5, 4), "sorted": bad == 0}), flush=True)
