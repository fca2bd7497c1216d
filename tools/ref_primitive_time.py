#!/usr/bin/env python
"""Times the reference's standalone scan / histogram studies (Docs/Snippets/PrefixSum-WorkEfficient.cu:229,
Docs/Snippets/Histogram.cu:35; unmodified, compiled into oracle/_ref) at their own configuration -- 2^24
elements, block size 512, host arrays in and out, everything inside the call as their main() times it -- and
checks the result.  Prints ONE JSON line.    python tools/ref_primitive_time.py --what scan|hist [--log2n 24]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def quiet_call(fn):
    sys.stdout.flush()
    saved = os.dup(1)
    rd, wr = os.pipe()
    os.dup2(wr, 1)
    t0 = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t0
    os.dup2(saved, 1)
    os.close(wr)
    chatter = os.read(rd, 1 << 16).decode(errors="replace")
    return out, dt, [ln for ln in chatter.splitlines() if ln.strip()][-4:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", choices=["scan", "hist"], required=True)
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--block", type=int, default=512)
    args = ap.parse_args()
    import torch
    import oracle as O
    name = "PrefixSum" if args.what == "scan" else "Histogram"
    if not O.ref_available(name):
        print(json.dumps({"unavailable": f"oracle/_ref/libref_{name.lower()}.so not built"}))
        return 0
    torch.zeros(1, device="cuda"); torch.cuda.synchronize()
    n = 1 << args.log2n
    rng = np.random.default_rng(1)
    best = None
    if args.what == "scan":
        x = rng.integers(0, 4, n, dtype=np.int32)                 # rand() & 0b11, PrefixSum-WorkEfficient.cu:402
        want = np.concatenate([[0], np.cumsum(x[:-1], dtype=np.int64)]).astype(np.int32)
        for impl, label in ((1, "BY_DEVICE"), (2, "BY_DEVICE_UNROLL2"), (3, "BY_DEVICE_UNROLL2_PAD")):
            for _ in range(2):
                out, dt, chat = quiet_call(lambda: O.ref_scan_by_device(x, impl, args.block))
                ok = bool(np.array_equal(out, want))
                if ok and (best is None or dt < best["ms"] * 1e-3):
                    best = {"variant": label, "ms": dt * 1e3, "reference_stdout": chat}
    else:
        bins = 16                                                    # numBits = 4, Histogram.cu:87
        x = rng.integers(0, bins, n, dtype=np.int32)
        want = np.bincount(x, minlength=bins).astype(np.int32)
        for _ in range(3):
            out, dt, chat = quiet_call(lambda: O.ref_histogram_by_device(x, bins, args.block))
            if np.array_equal(out, want) and (best is None or dt < best["ms"] * 1e-3):
                best = {"variant": "BY_DEVICE, 16 bins", "ms": dt * 1e3, "reference_stdout": chat}
    if best is None:
        print(json.dumps({"unavailable": "reference result incorrect"}))
        return 0
    best.update({"impl": f"reference Docs/Snippets {name} (unmodified, oracle/_ref)", "n": n, "block_size": args.block,
                 "elements_per_s": n / (best["ms"] * 1e-3),
                 "scope": "host arrays: cudaMalloc + H2D + kernels (+ host scan of block sums) + D2H inside, "
                          "as the snippet's own timer brackets it; best of the repetitions"})
    print(json.dumps(best))
    return 0


if __name__ == "__main__":
    sys.exit(main())
