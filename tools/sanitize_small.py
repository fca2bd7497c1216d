#!/usr/bin/env python
"""Small end-to-end exercise of every kernel, for compute-sanitizer (ONE tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Checks results against numpy so a sanitizer-clean run is also a correct run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cuda.radixsort_b200 as rs  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def main():
    rng = np.random.default_rng(7)
    small = int(os.environ.get("SANITIZE_N", "40001"))
    for variant in (-1, 0, 16):
        rs.set_param("variant", variant)
        for nbits in (8, 5, 1, 11):
            n = small if nbits != 1 else 3001
            k = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
            assert np.array_equal(host(rs.sort_keys(dev(k), nbits)), np.sort(k)), (variant, nbits)
        k = (rng.integers(0, 1 << 32, small, dtype=np.uint64).astype(np.uint32)) & 0x3FF
        v = np.arange(small, dtype=np.uint32)
        ko, vo = rs.sort_pairs(dev(k), dev(v), 8)
        idx = np.argsort(k, kind="stable")
        assert np.array_equal(host(ko), k[idx]) and np.array_equal(host(vo), v[idx]), variant
        # clustered inputs: run-length path
        k = np.repeat(rng.integers(0, 1 << 32, small // 50 + 1, dtype=np.uint64).astype(np.uint32), 50)[:small]
        assert np.array_equal(host(rs.sort_keys(dev(k), 8)), np.sort(k)), variant
    rs.set_param("variant", -1)
    rs.set_param("portion_tiles", 2)
    k = rng.integers(0, 1 << 32, 60001, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(host(rs.sort_keys(dev(k), 8)), np.sort(k))
    rs.set_param("portion_tiles", 0)
    h = rs.histogram(dev(k), 24, 8).cpu().numpy().view(np.uint32)
    assert np.array_equal(h, np.bincount(k >> 24, minlength=256).astype(np.uint32))
    p = rs.digit_pass(dev(k), 29, 3)
    assert np.array_equal(host(p), k[np.argsort(k >> 29, kind="stable")])
    x = (k >> 12)
    ref = np.concatenate([np.zeros(1, np.uint64), np.cumsum(x[:-1], dtype=np.uint64)]) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(host(rs.exclusive_scan(dev(x))), ref.astype(np.uint32))
    out = np.zeros_like(k)
    rs.sortByDevice(k, k.size, out, 8, 512)
    assert np.array_equal(out, np.sort(k))
    g = rs.generate("uniform", 5000)
    bad, *_ = rs.verify(rs.sort_keys(g, 8))
    assert bad == 0
    torch.cuda.synchronize()
    print("sanitize_small: all results correct")


if __name__ == "__main__":
    main()
