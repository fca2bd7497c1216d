"""CPU: pins the oracle (oracle/radix_oracle.c) to the reference before anything trusts it.

Sources of truth, strongest first:
  1. the two known-answer vectors of the reference's own test procedure (glibc rand(), unseeded;
     SourceCode/Baseline1.cu:140-160) with the fingerprints recorded in SURVEY.md section 8c;
  2. the committed golden fixtures produced by the UNMODIFIED reference (tests/golden/);
  3. the compiled reference itself (oracle/_ref), when it is present.
"""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_kat_debug_config(oracle):
    # #define DEBUG 1: n = 513, key = rand() & 0xFF, nBits = 4 (Baseline1.cu:141-142,155,174)
    k = oracle.glibc_rand_keys(513, 0xFF)
    assert list(k[:4]) == [103, 198, 105, 115]
    assert oracle.fnv1a64(k) == 0x0600861866D4BE5C
    for nbits in (4, 8):
        out = oracle.sort_keys(k, nbits)
        assert oracle.fnv1a64(out) == 0x714658018BFDCBDC
        assert (out[0], out[256], out[512]) == (0, 128, 255)


def test_kat_default_config(oracle):
    # default build: n = 2^24+1, key = rand(), nBits = 8 in main() / 4 as sort()'s default
    n = (1 << 24) + 1
    k = oracle.glibc_rand_keys(n)
    assert (k[0], k[1], k[-1], k.max()) == (1804289383, 846930886, 1923513432, 2147483611)
    assert oracle.fnv1a64(k) == 0x651B3F32FD6DA7FE
    for nbits in (4, 8):
        out = oracle.sort_keys(k, nbits)
        assert oracle.fnv1a64(out) == 0xE354BCFF33580302
        assert (out[0], out[n // 2], out[-1]) == (37, 1073726730, 2147483611)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(oracle, path):
    g = np.load(path)
    nbits = int(g["nbits"])
    assert np.array_equal(oracle.sort_keys(g["keys"], nbits), g["out"])
    if "out_parallel_algorithm_b512" in g:
        # the reference's CPU restatement of its GPU algorithm agrees with its sortByHost
        assert np.array_equal(g["out_parallel_algorithm_b512"], g["out"])


def test_golden_present():
    assert len(GOLDEN) >= 8


def test_against_compiled_reference(oracle):
    if not oracle.ref_available("Baseline1"):
        pytest.skip("oracle/_ref not built here")
    for kind, n, nbits in [("uniform", 100003, 8), ("uniform", 50000, 4), ("zipf", 30000, 7),
                           ("unique16", 20000, 3), ("uniform", 9999, 13), ("uniform", 4097, 16),
                           ("all_equal", 1000, 8), ("sorted", 5000, 6), ("reversed", 5000, 1)]:
        k = oracle.generate(kind, n)
        assert np.array_equal(oracle.sort_keys(k, nbits), oracle.ref_sort_by_host(k, nbits)), (kind, n, nbits)


@pytest.mark.parametrize("nbits", list(range(1, 17)))
def test_every_digit_width_sorts(oracle, nbits):
    k = oracle.generate("uniform", 20011, first=nbits)
    assert np.array_equal(oracle.sort_keys(k, nbits), np.sort(k))
    assert oracle.num_passes(nbits) == -(-32 // nbits)


def test_pairs_are_stable(oracle):
    # parity for pairs is unpinned by the reference (it has no key/value path): the oracle is
    # cross-checked against numpy's stable argsort instead
    for nbits in (1, 4, 5, 8, 11):
        k = oracle.generate("uniform", 30011) & 0x1FF
        v = oracle.generate("uniform", 30011, first=12345)
        ko, vo = oracle.sort_pairs(k, v, nbits)
        idx = np.argsort(k, kind="stable")
        assert np.array_equal(ko, k[idx]) and np.array_equal(vo, v[idx])


def test_tile_table_matches_definition(oracle):
    # Baseline4.cu:103-138: per-tile histogram + bin-major exclusive scan
    k = oracle.generate("uniform", 5000)
    tile, shift, nbits = 1024, 8, 4
    table, scan = oracle.tile_table(k, tile, shift, nbits)
    d = (k >> shift) & 15
    for t in range(table.shape[0]):
        assert np.array_equal(table[t], np.bincount(d[t * tile:(t + 1) * tile], minlength=16))
    flat = table.T.reshape(-1)  # bin-major
    excl = np.concatenate([[0], np.cumsum(flat)[:-1]]).reshape(16, -1).T
    assert np.array_equal(scan, excl.astype(np.uint32))
    # scan[t][d] is where the reference scatter (Baseline4.cu:234-241) puts the tile's first key of bin d
    out = np.empty_like(k)
    cursor = scan.astype(np.int64).copy()
    for i, key in enumerate(k):
        out[cursor[i // tile, d[i]]] = key
        cursor[i // tile, d[i]] += 1
    assert np.array_equal((out >> shift) & 15, np.sort(d))


def test_edge_sizes(oracle):
    for n in (0, 1, 2, 31, 32, 33):
        k = oracle.generate("uniform", n)
        assert np.array_equal(oracle.sort_keys(k, 8), np.sort(k))
    with pytest.raises(ValueError):
        oracle.sort_keys(np.zeros(4, np.uint32), 0)
    with pytest.raises(ValueError):
        oracle.sort_keys(np.zeros(4, np.uint32), 17)


def test_generators_are_deterministic_and_windowed(oracle):
    for kind in oracle.GEN_KINDS:
        whole = oracle.generate(kind, 1000, first=0, total=1000)
        part = oracle.generate(kind, 300, first=500, total=1000)
        assert np.array_equal(whole[500:800], part), kind
    assert oracle.is_sorted(oracle.generate("sorted", 4096))
    assert len(np.unique(oracle.generate("unique16", 10000))) == 16
    assert oracle.generate("uniform", 100000).max() > 0x80000000  # bit 31 is exercised
    z = oracle.generate("zipf", 200000)
    vals, counts = np.unique(z, return_counts=True)
    assert counts.max() > 0.05 * z.size and len(vals) > 5000  # heavy head, long tail


def test_multiset_fingerprint_is_order_independent(oracle):
    k = oracle.generate("uniform", 10000)
    assert oracle.multiset_fingerprint(k) == oracle.multiset_fingerprint(np.sort(k))
    k2 = k.copy()
    k2[5] ^= 1
    assert oracle.multiset_fingerprint(k) != oracle.multiset_fingerprint(k2)
