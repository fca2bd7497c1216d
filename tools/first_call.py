#!/usr/bin/env python
"""First-call cost of the host-pointer entry point (the sortByDevice replacement) in a fresh process:
    python tools/first_call.py [--log2n 28] [--warmup] [--pageable]
Prints the wall time of the first, second and third call (and of b200sort_warmup when asked)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--warmup", action="store_true")
    ap.add_argument("--pageable", action="store_true")
    args = ap.parse_args()
    import torch
    import cuda.radixsort_b200 as rs
    rs.load()
    n = 1 << args.log2n
    t0 = time.perf_counter()
    torch.zeros(1, device="cuda"); torch.cuda.synchronize()     # CUDA context (the reference's main() does cudaFree(0))
    ctx_ms = (time.perf_counter() - t0) * 1e3
    rng = np.random.default_rng(7)
    k = rng.integers(0, 1 << 32, n, dtype=np.uint32)
    if args.pageable:
        a_in, a_out = k, np.empty_like(k)
    else:
        h_in = torch.empty(n, dtype=torch.int32).pin_memory(); h_out = torch.empty(n, dtype=torch.int32).pin_memory()
        a_in, a_out = h_in.numpy().view(np.uint32), h_out.numpy().view(np.uint32)
        a_in[:] = k
    out = {"n": n, "host_buffers": "pageable" if args.pageable else "pinned", "cuda_context_ms": round(ctx_ms, 1)}
    if args.warmup:
        t0 = time.perf_counter(); rs.warmup(n); out["warmup_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    calls = []
    for _ in range(3):
        t0 = time.perf_counter(); rs.sort(a_in, n, a_out, rs.SORT_BY_DEVICE, 8, 512); calls.append(round((time.perf_counter() - t0) * 1e3, 1))
    assert a_out[0] <= a_out[n // 2] <= a_out[-1]
    out["call_ms"] = calls
    print(json.dumps(out))


if __name__ == "__main__":
    main()
