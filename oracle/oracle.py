"""ctypes/numpy front end of oracle/liboracle.so and of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY (see oracle/radix_oracle.c).  Function -> reference citation:

  sort_keys / sort_pairs   SourceCode/Baseline1.cu:15-64 (sortByHost), restated in C
  tile_table               SourceCode/Baseline4.cu:103-138 (tile histogram + bin-major scan)
  glibc_rand_keys          SourceCode/Baseline1.cu:152-158 (the reference's test input)
  ref_sort_by_host         the reference's own sortByHost, unmodified, from oracle/_ref
  ref_sort_by_device       the reference's own Parallel7 sortByDevice (needs a GPU)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle as _build

__all__ = [
    "GEN_KINDS", "fnv1a64", "generate", "glibc_rand_keys", "is_sorted", "multiset_fingerprint",
    "num_passes", "ref_available", "ref_sort_by_device", "ref_sort_by_host", "ref_scan_by_device", "ref_histogram_by_device",
    "ref_sort_by_host_parallel_algorithm", "sort_keys", "sort_pairs", "tile_table", "zipf_cdf",
]

_U32P = C.POINTER(C.c_uint32)
_lib = None
_ref_libs: dict[str, C.CDLL] = {}

GEN_KINDS = {"uniform": 0, "zipf": 1, "unique16": 2, "all_equal": 3, "sorted": 4,
             "reversed": 5, "iota": 6}


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.build_oracle()
        lib = C.CDLL(path)
        lib.oracle_num_passes.argtypes = [C.c_int]
        lib.oracle_num_passes.restype = C.c_int
        lib.oracle_sort_keys.argtypes = [_U32P, C.c_int64, _U32P, C.c_int]
        lib.oracle_sort_keys.restype = C.c_int
        lib.oracle_sort_pairs.argtypes = [_U32P, _U32P, C.c_int64, _U32P, _U32P, C.c_int]
        lib.oracle_sort_pairs.restype = C.c_int
        lib.oracle_tile_table.argtypes = [_U32P, C.c_int64, C.c_int64, C.c_int, C.c_int, _U32P, _U32P]
        lib.oracle_tile_table.restype = C.c_int64
        lib.oracle_fnv1a64_words.argtypes = [_U32P, C.c_int64]
        lib.oracle_fnv1a64_words.restype = C.c_uint64
        lib.oracle_fill_glibc_rand.argtypes = [_U32P, C.c_int64, C.c_uint32]
        lib.oracle_fill_glibc_rand.restype = None
        lib.oracle_zipf_cdf.argtypes = [_U32P]
        lib.oracle_zipf_cdf.restype = None
        lib.oracle_generate.argtypes = [_U32P, C.c_int64, C.c_int64, C.c_int, C.c_int64, _U32P]
        lib.oracle_generate.restype = C.c_int
        lib.oracle_multiset_fingerprint.argtypes = [_U32P, C.c_int64, C.POINTER(C.c_uint64)]
        lib.oracle_multiset_fingerprint.restype = None
        lib.oracle_is_sorted.argtypes = [_U32P, C.c_int64]
        lib.oracle_is_sorted.restype = C.c_int
        _lib = lib
    return _lib


def _u32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(_U32P)


def num_passes(nbits: int) -> int:
    return _load().oracle_num_passes(nbits)


def sort_keys(keys, nbits: int = 8) -> np.ndarray:
    k = _u32(keys)
    out = np.empty_like(k)
    if _load().oracle_sort_keys(_p(k), k.size, _p(out), nbits) != 0:
        raise ValueError(f"oracle_sort_keys rejected nbits={nbits}")
    return out


def sort_pairs(keys, vals, nbits: int = 8):
    k, v = _u32(keys), _u32(vals)
    assert k.size == v.size
    ko, vo = np.empty_like(k), np.empty_like(v)
    if _load().oracle_sort_pairs(_p(k), _p(v), k.size, _p(ko), _p(vo), nbits) != 0:
        raise ValueError(f"oracle_sort_pairs rejected nbits={nbits}")
    return ko, vo


def tile_table(keys, tile: int, shift: int, nbits: int):
    k = _u32(keys)
    bins = 1 << nbits
    tiles = (k.size + tile - 1) // tile
    table = np.zeros((tiles, bins), dtype=np.uint32)
    scan = np.zeros((tiles, bins), dtype=np.uint32)
    got = _load().oracle_tile_table(_p(k), k.size, tile, shift, nbits, _p(table), _p(scan))
    assert got == tiles
    return table, scan


def fnv1a64(words) -> int:
    w = _u32(words)
    return int(_load().oracle_fnv1a64_words(_p(w), w.size))


def glibc_rand_keys(n: int, and_mask: int = 0xFFFFFFFF) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    _load().oracle_fill_glibc_rand(_p(out), n, and_mask)
    return out


_cdf = None


def zipf_cdf() -> np.ndarray:
    global _cdf
    if _cdf is None:
        c = np.empty(65536, dtype=np.uint32)
        _load().oracle_zipf_cdf(_p(c))
        _cdf = c
    return _cdf


def generate(kind: str, count: int, first: int = 0, total: int | None = None) -> np.ndarray:
    """The synthetic workloads of SURVEY.md 8d (same bytes as the device generators)."""
    out = np.empty(count, dtype=np.uint32)
    cdf = zipf_cdf() if kind == "zipf" else None
    rc = _load().oracle_generate(_p(out), first, count, GEN_KINDS[kind],
                                 total if total is not None else first + count,
                                 _p(cdf) if cdf is not None else None)
    if rc != 0:
        raise ValueError(kind)
    return out


def multiset_fingerprint(keys) -> tuple[int, int, int]:
    k = _u32(keys)
    out = (C.c_uint64 * 3)()
    _load().oracle_multiset_fingerprint(_p(k), k.size, out)
    return int(out[0]), int(out[1]), int(out[2])


def is_sorted(keys) -> bool:
    k = _u32(keys)
    return bool(_load().oracle_is_sorted(_p(k), k.size))


# --------------------------------------------------------------------------------------
# The compiled reference itself (oracle/_ref), when it has been built.

def _ref(name: str) -> C.CDLL | None:
    if name not in _ref_libs:
        path = _build.ref_so(name)
        if not os.path.exists(path):
            _build.build_reference(files=(name,))
        if not os.path.exists(path):
            return None
        _ref_libs[name] = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
    return _ref_libs[name]


def ref_available(name: str = "Baseline1") -> bool:
    try:
        return _ref(name) is not None
    except OSError:
        return False


def ref_sort_by_host(keys, nbits: int = 8, name: str = "Baseline1") -> np.ndarray:
    """void sortByHost(const uint32_t*, int, uint32_t*, int) -- SourceCode/Baseline1.cu:15."""
    lib = _ref(name)
    if lib is None:
        raise RuntimeError("oracle/_ref not built (no /root/reference here)")
    fn = lib._Z10sortByHostPKjiPji
    fn.argtypes = [_U32P, C.c_int, _U32P, C.c_int]
    fn.restype = None
    k = _u32(keys)
    out = np.empty_like(k)
    fn(_p(k), k.size, _p(out), nbits)
    return out


def ref_sort_by_host_parallel_algorithm(keys, nbits: int, block_size: int) -> np.ndarray:
    """sortByHostUsingParallelAlgorithm -- SourceCode/Baseline4.cu:67 (needs 32 % nbits == 0)."""
    assert 32 % nbits == 0
    lib = _ref("Baseline4")
    if lib is None:
        raise RuntimeError("oracle/_ref not built")
    fn = lib._Z32sortByHostUsingParallelAlgorithmPKjiPjii
    fn.argtypes = [_U32P, C.c_int, _U32P, C.c_int, C.c_int]
    fn.restype = None
    k = _u32(keys)
    out = np.empty_like(k)
    fn(_p(k), k.size, _p(out), nbits, block_size)
    return out


def ref_sort_by_device(keys, nbits: int = 8, block_size: int = 512) -> np.ndarray:
    """The reference's best GPU version, Parallel7 sortByDevice (SourceCode/Parallel7.cu:530).

    Needs a GPU.  Parallel7 keeps function-static device buffers sized by its first call
    (Parallel7.cu:203-218), so within one process call it with non-increasing n only.
    """
    lib = _ref("Parallel7")
    if lib is None:
        raise RuntimeError("oracle/_ref not built")
    fn = lib._Z12sortByDevicePKjiPjii
    fn.argtypes = [_U32P, C.c_int, _U32P, C.c_int, C.c_int]
    fn.restype = None
    k = _u32(keys)
    out = np.empty_like(k)
    fn(_p(k), k.size, _p(out), nbits, block_size)
    return out


# --------------------------------------------------------------------------------------
# The reference's standalone kernel studies (Docs/Snippets), compiled where they lie: same-box bars for the
# scan and histogram primitives (bench.py --workload scan|hist).  Both need a GPU and print their own "Time:".

def ref_scan_by_device(values, implementation: int = 2, block_size: int = 512) -> np.ndarray:
    """void scanByDevice(const int*, int, int*, Implementation, int) -- Docs/Snippets/PrefixSum-WorkEfficient.cu:229
    (exclusive scan; implementation 1 = BY_DEVICE, 2 = BY_DEVICE_UNROLL2, 3 = BY_DEVICE_UNROLL2_PAD; host arrays,
    cudaMalloc + H2D + kernels + host scan of the block sums + D2H inside)."""
    lib = _ref("PrefixSum")
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_prefixsum.so not built")
    fn = lib._Z12scanByDevicePKiiPi14Implementationi
    i32p = C.POINTER(C.c_int32)
    fn.argtypes = [i32p, C.c_int, i32p, C.c_int, C.c_int]
    fn.restype = None
    x = np.ascontiguousarray(values, dtype=np.int32)
    out = np.empty_like(x)
    fn(x.ctypes.data_as(i32p), x.size, out.ctypes.data_as(i32p), implementation, block_size)
    return out


def ref_histogram_by_device(values, num_bins: int, block_size: int = 512) -> np.ndarray:
    """void histogram(const int*, int, int*, int numBins, Implementation = BY_DEVICE, int blockSize) --
    Docs/Snippets/Histogram.cu:35 (values must be < numBins; host arrays, everything inside)."""
    lib = _ref("Histogram")
    if lib is None:
        raise RuntimeError("oracle/_ref/libref_histogram.so not built")
    fn = lib._Z9histogramPKiiPii14Implementationi
    i32p = C.POINTER(C.c_int32)
    fn.argtypes = [i32p, C.c_int, i32p, C.c_int, C.c_int, C.c_int]
    fn.restype = None
    x = np.ascontiguousarray(values, dtype=np.int32)
    out = np.zeros(num_bins, dtype=np.int32)
    fn(x.ctypes.data_as(i32p), x.size, out.ctypes.data_as(i32p), num_bins, 1, block_size)
    return out
