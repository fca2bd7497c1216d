#!/usr/bin/env python
"""Times the UNMODIFIED reference GPU sort -- Parallel7 sortByDevice(h_in, n, h_out, numBits, blockSize),
/root/reference/SourceCode/Parallel7.cu:530-639, compiled into oracle/_ref/libref_parallel7.so -- on this
box's GPU, the way the reference's main() calls it (pageable host arrays; cudaMalloc/H2D/D2H inside, 88
kernels with a device synchronisation after each).  One process per n: Parallel7 keeps function-static
device buffers sized by its first call (Parallel7.cu:203-218).  Prints ONE JSON line.

    python tools/ref_gpu_time.py [--log2n 28] [--nbits 8] [--block 512]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--nbits", type=int, default=8)
    ap.add_argument("--block", type=int, default=512)
    args = ap.parse_args()
    import torch
    import oracle as O
    if not O.ref_available("Parallel7"):
        print(json.dumps({"unavailable": "oracle/_ref/libref_parallel7.so not built"}))
        return 0
    torch.zeros(1, device="cuda")          # create the CUDA context outside the timed call (main() does cudaFree(0))
    torch.cuda.synchronize()
    n = 1 << args.log2n
    keys = O.generate("uniform", n)
    # redirect the reference's own printf lines (stage times) away from our JSON
    sys.stdout.flush()
    saved = os.dup(1)
    rd, wr = os.pipe()
    os.dup2(wr, 1)
    t0 = time.perf_counter()
    out = O.ref_sort_by_device(keys, args.nbits, args.block)
    dt = time.perf_counter() - t0
    os.dup2(saved, 1)
    os.close(wr)
    chatter = os.read(rd, 1 << 16).decode(errors="replace")
    ok = bool(np.all(out[:-1] <= out[1:])) and int(out[0]) == int(keys.min()) and int(out[-1]) == int(keys.max())
    print(json.dumps({"impl": "reference Parallel7 sortByDevice (unmodified, oracle/_ref)", "n": n, "nbits": args.nbits,
                      "block_size": args.block, "ms": dt * 1e3, "keys_per_s": n / dt, "sorted": ok,
                      "host_buffers": "pageable (numpy)", "first_call_in_process": True,
                      "reference_stdout": [ln for ln in chatter.splitlines() if ln.strip()][-8:]}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
