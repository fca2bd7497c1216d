#!/usr/bin/env python
"""Times every digit-pass kernel variant on the bench workload (device-resident, CUDA events) and
checks each result (sortedness + multiset fingerprint).  One JSON line per variant.

    python tools/sweep.py [--log2n 28] [--steps 5] [--variants 0,1,2] [--workload uniform|pairs|zipf|...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import cuda.radixsort_b200 as rs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--nbits", type=int, default=8)
    ap.add_argument("--variants", default="all")
    ap.add_argument("--workload", default="uniform")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    rs.load()
    n = 1 << args.log2n
    pairs = args.workload == "pairs"
    dist = "uniform" if pairs else args.workload
    cdf = None
    if dist == "zipf":
        import oracle as O
        cdf = O.zipf_cdf()
    keys = rs.generate(dist, n, zipf_cdf=cdf)
    vals = torch.arange(n, dtype=torch.int32, device="cuda") if pairs else None
    out = torch.empty_like(keys)
    vout = torch.empty_like(keys) if pairs else None
    ws = rs.Workspace("cuda")
    ws.get(rs.temp_bytes(n, args.nbits, pairs))
    _, s0, h0, x0 = rs.verify(keys)
    nv = rs.get_param("num_variants")
    variants = range(nv) if args.variants == "all" else [int(v) for v in args.variants.split(",")]
    print(json.dumps({"atomic_rank_ok": rs.get_param("atomic_rank_ok"), "n": n, "workload": args.workload,
                      "nbits": args.nbits}), flush=True)
    for v in variants:
        rs.set_param("variant", v)
        eff = rs.get_param("effective_variant")
        if eff != v and not rs.get_param("tuning_build"):
            continue  # not in the product build: B200_TUNING=1 python -m cuda.radixsort_b200.build --force

        def step():
            if pairs:
                rs.sort_pairs(keys, vals, args.nbits, out_keys=out, out_vals=vout, workspace=ws)
            else:
                rs.sort_keys(keys, args.nbits, out=out, workspace=ws)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        rs.profile_enable(True)
        rs.profile_read()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        b.record()
        torch.cuda.synchronize()
        prof = rs.profile_read()
        rs.profile_enable(False)
        ms = a.elapsed_time(b) / args.steps
        bad, s1, h1, x1 = rs.verify(out)
        ok = bad == 0 and (s1, h1, x1) == (s0, h0, x0)
        if pairs and ok:
            # stability/permutation check: keys[vout] == out
            ok = bool(torch.equal(keys[vout.long()], out))
        hist = [m for t, m in prof if t == 0]
        passes = [m for t, m in prof if t >= 1]
        print(json.dumps({"variant": v, "effective": eff, "tile": rs.tile_keys(pairs), "ms": round(ms, 4),
                          "gkeys_s": round(n / ms / 1e6, 2), "hist_ms": round(sum(hist) / max(1, len(hist)), 4),
                          "pass_ms": round(sum(passes) / max(1, len(passes)), 4), "ok": ok}), flush=True)
    rs.set_param("variant", -1)


if __name__ == "__main__":
    main()
