"""Times the single-process multi-GPU host entry point (b200sort_mgpu_keys_host) on pinned and
pageable host arrays: `python tools/mgpu_host_bench.py --log2n 28 --devices 0,1 [--pageable]`.
Prints one JSON line per configuration: wall time of the blocking call (host arrays in, host
arrays out), the device-event phase times the library reports, and a sortedness + checksum check.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def host_array(n, pinned):
    import torch
    t = torch.empty(n, dtype=torch.int32)
    if pinned:
        t = t.pin_memory()
    return t, t.numpy().view(np.uint32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--pageable", action="store_true")
    ap.add_argument("--pairs", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nbits", type=int, default=8)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build(reference=False)
    import cuda.radixsort_b200 as rs

    n = 1 << a.log2n
    devices = [int(d) for d in a.devices.split(",")]
    keep_in, k = host_array(n, not a.pageable)
    keep_out, out = host_array(n, not a.pageable)
    rng = np.random.default_rng(7)
    step = 1 << 24
    block = rng.integers(0, 1 << 32, min(step, n), dtype=np.uint64).astype(np.uint32)
    for i, lo in enumerate(range(0, n, step)):   # one random block, re-keyed per block (cheap at 2^32)
        np.bitwise_xor(block[:min(step, n - lo)], np.uint32((i * 0x9E3779B1) & 0xFFFFFFFF), out=k[lo:lo + step])
    if a.pairs:
        keep_v, v = host_array(n, not a.pageable)
        keep_ov, ov = host_array(n, not a.pageable)
        v[:] = np.arange(n, dtype=np.uint32)

    def call():
        if a.pairs:
            rs.sort_pairs_by_devices(k, v, n, out, ov, a.nbits, 512, devices)
        else:
            rs.sort_by_devices(k, n, out, a.nbits, 512, devices)

    call()  # allocates the cached buffers
    call()
    times, stats = [], None
    for _ in range(a.reps):
        t0 = time.perf_counter()
        call()
        times.append((time.perf_counter() - t0) * 1e3)
        stats = rs.mgpu_last_stats()
    ok = int(out.sum(dtype=np.uint64)) == int(k.sum(dtype=np.uint64))
    for lo in range(0, n, 1 << 28):
        hi = min(n, lo + (1 << 28) + 1)
        ok = ok and bool(np.all(out[lo + 1:hi] >= out[lo:hi - 1]))
    if a.pairs:
        ok = ok and bool(np.array_equal(k[ov[:: max(1, n // 4096)]], out[:: max(1, n // 4096)]))
    ms = float(np.median(times))
    print(json.dumps({"tool": "mgpu_host_bench", "n": n, "devices": devices, "pinned": not a.pageable,
                      "pairs": a.pairs, "nbits": a.nbits, "ms": round(ms, 3), "ms_min": round(min(times), 3),
                      "gkeys_per_s": round(n / ms / 1e6, 3), "correct": ok,
                      "phases": {kk: round(vv, 3) for kk, vv in stats.items()}}), flush=True)
    rs.shutdown()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
