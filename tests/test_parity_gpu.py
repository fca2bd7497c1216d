"""GPU parity: the CUDA path, called through the C ABI (ctypes -> libb200sort.so), against the
oracle on the same inputs.  Integer work: every comparison is bit-exact.

Mirrors the reference's own test procedure (SourceCode/Parallel7.cu:704-767: sort on host,
sort on device, compare element by element) and closes the gaps SURVEY.md section 4 lists
(bit 31, duplicates, tiny and ragged n, digit widths that do not divide 32, n = 0)."""
import glob
import os

import numpy as np
import pytest

from conftest import to_dev, to_host

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def dev_sort(rs, keys, nbits):
    return to_host(rs.sort_keys(to_dev(keys), nbits))


# ---------------------------------------------------------------------------- known answers

def test_kat_debug_config_host_entry(rs, oracle):
    k = oracle.glibc_rand_keys(513, 0xFF)
    out = np.zeros_like(k)
    rs.sort(k, k.size, out, rs.SORT_BY_DEVICE, 4, 512)       # the reference's call, Parallel7.cu:763
    assert oracle.fnv1a64(out) == 0x714658018BFDCBDC
    assert np.array_equal(out, oracle.sort_keys(k, 4))


@pytest.mark.parametrize("nbits", [8, 4])
def test_kat_default_config(rs, oracle, nbits):
    n = (1 << 24) + 1
    k = oracle.glibc_rand_keys(n)
    out = np.zeros_like(k)
    rs.sortByDevice(k, n, out, nbits, 512)
    assert oracle.fnv1a64(out) == 0xE354BCFF33580302
    assert (out[0], out[n // 2], out[-1]) == (37, 1073726730, 2147483611)
    if nbits == 8:
        assert np.array_equal(out, oracle.sort_keys(k, nbits))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(rs, path):
    g = np.load(path)
    assert np.array_equal(dev_sort(rs, g["keys"], int(g["nbits"])), g["out"])


# ---------------------------------------------------------------------------- sizes, widths

def edge_sizes(rs):
    t = rs.tile_keys(False)
    return [0, 1, 2, 31, 32, 33, 255, 256, 257, t - 1, t, t + 1, 2 * t, 2 * t + 1, 3 * t - 1, (1 << 20) + 1]


def test_edge_sizes(rs, oracle):
    for n in edge_sizes(rs):
        k = oracle.generate("uniform", n, first=n)
        assert np.array_equal(dev_sort(rs, k, 8), oracle.sort_keys(k, 8)), n
        if n:
            out = np.zeros_like(k)
            rs.sortByDevice(k, n, out, 8, 512)
            assert np.array_equal(out, np.sort(k)), n


@pytest.mark.parametrize("nbits", list(range(1, 17)))
def test_every_digit_width(rs, oracle, nbits):
    n = 50021 if nbits > 2 else 20011
    k = oracle.generate("uniform", n, first=nbits * 1000)
    assert np.array_equal(dev_sort(rs, k, nbits), oracle.sort_keys(k, nbits))


@pytest.mark.parametrize("kind", ["uniform", "zipf", "unique16", "all_equal", "sorted", "reversed", "iota"])
@pytest.mark.parametrize("nbits", [8, 4])
def test_distributions(rs, oracle, kind, nbits):
    n = (1 << 21) + 77
    k = oracle.generate(kind, n)
    assert np.array_equal(dev_sort(rs, k, nbits), oracle.sort_keys(k, nbits))


def test_host_entry_staging_chunk_boundaries(rs, oracle):
    """Pageable host arrays go through 32 MiB pinned staging chunks: sizes around the chunk
    boundaries, including a last chunk of a single key (the reference's own n = 2^24 + 1)."""
    chunk = (32 << 20) // 4
    for n in (2 * chunk - 1, 2 * chunk, 2 * chunk + 1, 3 * chunk + 5, chunk // 4 + 1):
        k = oracle.generate("uniform", n, first=n)
        out = np.zeros_like(k)
        rs.sortByDevice(k, n, out, 8, 512)
        assert np.array_equal(out, np.sort(k)), n
    n = 2 * chunk + 1
    k = oracle.generate("uniform", n) & 0xFFFF
    v = np.arange(n, dtype=np.uint32)
    hk, hv = np.zeros_like(k), np.zeros_like(v)
    rs.sort_pairs_by_device(k, v, n, hk, hv, 8, 512)
    idx = np.argsort(k, kind="stable")
    assert np.array_equal(hk, k[idx]) and np.array_equal(hv, v[idx])


@pytest.mark.parametrize("pinned", [False, True])
def test_host_entry_overlapped_path(rs, oracle, pinned):
    """From 2^24 keys on, the host-pointer entry (the sortByDevice replacement) uploads in chunks with the
    top-4-bit histogram accumulated per chunk, splits on those bits (MSD) and sorts / downloads bucket by
    bucket: same answer as the single-shot path and as the oracle -- uniform keys, keys that all fall into ONE
    bucket, an empty top bucket, and stable pairs; pageable and pinned arrays."""
    import torch
    assert rs.get_param("host_overlap") == 1
    n = (1 << 24) + 4099

    def buf(a):
        if not pinned:
            return a
        t = torch.from_numpy(a.view(np.int32)).pin_memory()
        return t.numpy().view(np.uint32)

    for kind, mask, orv in (("uniform", 0xFFFFFFFF, 0), ("uniform", 0x0FFFFFFF, 0x30000000), ("uniform", 0x7FFFFFFF, 0),
                            ("zipf", 0xFFFFFFFF, 0)):
        k = (oracle.generate(kind, n) & np.uint32(mask)) | np.uint32(orv)
        want = oracle.sort_keys(k, 8)
        hin, out = buf(k.copy()), buf(np.zeros_like(k))
        rs.sortByDevice(hin, n, out, 8, 512)
        assert np.array_equal(out, want), (kind, hex(mask))
        assert np.array_equal(hin, k)                        # the input is not modified
    rs.set_param("host_overlap", 0)
    try:
        out2 = np.zeros_like(k)
        rs.sortByDevice(k, n, out2, 8, 512)                  # single-shot path
    finally:
        rs.set_param("host_overlap", 1)
    assert np.array_equal(out2, want)
    kk = oracle.generate("uniform", n) & np.uint32(0xF00000FF)   # 16 buckets x 256 values: long runs of equal keys
    v = np.arange(n, dtype=np.uint32)
    hk, hv = buf(np.zeros_like(kk)), buf(np.zeros_like(v))
    rs.sort_pairs_by_device(buf(kk.copy()), buf(v.copy()), n, hk, hv, 8, 512)
    rk, rv = oracle.sort_pairs(kk, v, 8)
    assert np.array_equal(hk, rk) and np.array_equal(hv, rv)


def test_keys_with_bit31_and_extremes(rs, oracle):
    k = oracle.generate("uniform", 100000)
    k[:5] = [0, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 0xFFFFFFFF]
    k[-3:] = [0xFFFFFFFF, 0, 0xFFFFFFFE]
    for nbits in (8, 5, 3):
        assert np.array_equal(dev_sort(rs, k, nbits), np.sort(k))
    # all-ones keys collide with the padding value of a ragged last tile
    k = np.full(rs.tile_keys(False) + 100, 0xFFFFFFFF, np.uint32)
    k[::7] = 5
    assert np.array_equal(dev_sort(rs, k, 8), np.sort(k))


def test_input_is_not_modified_and_unaligned_views(rs, oracle):
    import torch
    k = oracle.generate("uniform", 300001)
    d = to_dev(k)
    for off in (0, 1, 2, 3):      # 4-byte aligned but not 16-byte aligned sub-arrays
        view = d[off:]
        out = rs.sort_keys(view, 8)
        assert np.array_equal(to_host(out), np.sort(k[off:]))
    assert np.array_equal(to_host(d), k)
    with pytest.raises(rs.RadixSortError):
        rs.sort_keys(d, 8, out=d)          # aliasing is rejected, not silently wrong
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------- pairs

@pytest.mark.parametrize("nbits", [8, 4, 5, 11])
def test_pairs_stable(rs, oracle, nbits):
    n = 400003
    k = oracle.generate("uniform", n) & 0xFFF           # many duplicates
    v = np.arange(n, dtype=np.uint32)
    ko, vo = rs.sort_pairs(to_dev(k), to_dev(v), nbits)
    rk, rv = oracle.sort_pairs(k, v, nbits)
    assert np.array_equal(to_host(ko), rk)
    assert np.array_equal(to_host(vo), rv)              # equal keys keep input order


def test_pairs_edge_sizes_and_host_entry(rs, oracle):
    t = rs.tile_keys(True)
    for n in (1, 2, 33, t - 1, t, t + 1, 2 * t + 5):
        k = oracle.generate("unique16", n)
        v = oracle.generate("uniform", n, first=99)
        rk, rv = oracle.sort_pairs(k, v, 8)
        ko, vo = rs.sort_pairs(to_dev(k), to_dev(v), 8)
        assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv), n
        hk, hv = np.zeros_like(k), np.zeros_like(v)
        rs.sort_pairs_by_device(k, v, n, hk, hv, 8, 512)
        assert np.array_equal(hk, rk) and np.array_equal(hv, rv), n


@pytest.mark.parametrize("pattern", ["sorted", "iota", "two_values", "runs_of_16", "lane_period_16", "sawtooth"])
def test_pairs_stable_on_clustered_keys(rs, oracle, pattern):
    """Inputs whose equal digits sit in neighbouring lanes (the run-length rank path) or in a
    fixed lane period (same-address atomics inside one warp instruction): stability must hold."""
    n = 700001
    i = np.arange(n, dtype=np.uint64)
    if pattern == "sorted":
        k = oracle.generate("sorted", n)
    elif pattern == "iota":
        k = oracle.generate("iota", n)
    elif pattern == "two_values":
        k = np.where(oracle.generate("uniform", n) & 1, 0xCAFEBABE, 0x12345678).astype(np.uint32)
    elif pattern == "runs_of_16":
        k = ((i // 16) * 2654435761 % (1 << 32)).astype(np.uint32)
    elif pattern == "lane_period_16":
        k = ((i % 16) * 0x01010101).astype(np.uint32)
    else:
        k = ((i % 1000) << 12).astype(np.uint32)
    v = np.arange(n, dtype=np.uint32)
    for nbits in (8, 4):
        ko, vo = rs.sort_pairs(to_dev(k), to_dev(v), nbits)
        rk, rv = oracle.sort_pairs(k, v, nbits)
        assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv), (pattern, nbits)
    assert np.array_equal(dev_sort(rs, k, 8), np.sort(k))


def test_pairs_all_equal_keys_is_identity_on_values(rs, oracle):
    n = 100000
    k = oracle.generate("all_equal", n)
    v = oracle.generate("uniform", n)
    ko, vo = rs.sort_pairs(to_dev(k), to_dev(v), 8)
    assert np.array_equal(to_host(vo), v) and np.array_equal(to_host(ko), k)


# ---------------------------------------------------------------------------- internals via the ABI

def test_multi_launch_portions(rs, oracle):
    """Force several launches per pass (the path n >= 2^30 takes) at a small n."""
    rs.set_param("portion_tiles", 3)
    try:
        for n in (3 * rs.tile_keys(False) * 4 + 11, 200001):
            k = oracle.generate("zipf", n)
            assert np.array_equal(dev_sort(rs, k, 8), np.sort(k)), n
            assert np.array_equal(dev_sort(rs, k, 5), np.sort(k)), n
        kk = oracle.generate("uniform", 150001) & 0xFF
        v = np.arange(kk.size, dtype=np.uint32)
        ko, vo = rs.sort_pairs(to_dev(kk), to_dev(v), 4)
        rk, rv = oracle.sort_pairs(kk, v, 4)
        assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv)
    finally:
        rs.set_param("portion_tiles", 0)


def _variants_in_this_build():
    """Every geometry of csrc/launch.h with the tuning build (B200_TUNING=1), else the kernels the product
    library carries for the 8-bit digit: the column sweep (36) and the default, the column sweep with two ranking
    chains (95)."""
    if os.environ.get("B200_TUNING"):
        try:
            import ctypes
            from cuda.radixsort_b200 import build as lib_build
            lib = ctypes.CDLL(lib_build.build())          # (re)builds with the tuning flags when needed
            if lib.b200sort_get_param(b"tuning_build") == 1:
                return list(range(lib.b200sort_get_param(b"num_variants")))
        except (OSError, RuntimeError):
            pass
    return [36, 95]


@pytest.mark.parametrize("variant", _variants_in_this_build())
def test_kernel_variants(rs, oracle, variant):
    rs.set_param("variant", variant)
    try:
        assert variant < rs.get_param("num_variants")
        if rs.get_param("effective_variant") != variant:
            # the product library carries only the kernels its automatic choice can select; the other
            # geometries need the tuning build (B200_TUNING=1 python -m cuda.radixsort_b200.build --force)
            assert rs.get_param("tuning_build") == 0 or rs.get_param("atomic_rank_ok") != 1
            pytest.skip("variant not in this build")
        n = (1 << 20) + 12345
        k = oracle.generate("uniform", n)
        assert np.array_equal(dev_sort(rs, k, 8), oracle.sort_keys(k, 8))
        z = oracle.generate("zipf", n)
        assert np.array_equal(dev_sort(rs, z, 8), oracle.sort_keys(z, 8))
        kk = k & 0x3FF
        v = np.arange(n, dtype=np.uint32)
        ko, vo = rs.sort_pairs(to_dev(kk), to_dev(v), 8)
        rk, rv = oracle.sort_pairs(kk, v, 8)
        assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv)
    finally:
        rs.set_param("variant", -1)


def test_atomic_rank_selftest_and_stability(rs, oracle):
    """RANK_ATOMIC relies on lane-ordered same-address shared atomics; the library only uses it
    after its on-device self test passed.  Whatever the verdict, variant 1 must stay stable."""
    verdict = rs.get_param("atomic_rank_ok")
    assert verdict in (0, 1)
    rs.set_param("variant", 1)
    try:
        if rs.get_param("tuning_build"):
            assert rs.get_param("effective_variant") == (1 if verdict == 1 else 36)   # else: the spec-safe column sweep
        else:   # the product library does not carry the atomic-rank kernels: the request gets the default kernel
            assert rs.get_param("effective_variant") == 95
        assert rs.get_param("atomic_rank_ok") == verdict
        for kind in ("unique16", "all_equal", "zipf"):
            n = 300007
            k = oracle.generate(kind, n)
            v = np.arange(n, dtype=np.uint32)
            for nbits in (8, 4, 1):
                ko, vo = rs.sort_pairs(to_dev(k), to_dev(v), nbits)
                rk, rv = oracle.sort_pairs(k, v, nbits)
                assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv), (kind, nbits)
    finally:
        rs.set_param("variant", -1)


@pytest.mark.parametrize("kernel", ["colsweep", "default"])
def test_safe_rank_mode_every_width_keys_pairs_and_destinations(rs, oracle, kernel):
    """safe_rank=1 confines the library to kernels whose ranking follows from the PTX memory model: the
    column-sweep kernel (lane-private counters, warp turns ordered by named barriers) replaces every
    atomic-rank variant, for every digit width, for pairs and for per-bin destinations.  The default kernel
    (the column sweep with two chains, variant 95, digits of >= 4 bits) is such a kernel itself and stays."""
    rs.set_param("safe_rank", 1)
    if kernel == "colsweep":
        # tuning build: an atomic-rank request, which safe_rank replaces by the column sweep (variant 36);
        # product build (no atomic-rank kernels): that kernel by its number
        rs.set_param("variant", 1 if rs.get_param("tuning_build") else 36)
    try:
        assert rs.get_param("safe_rank") == 1
        if kernel == "colsweep":
            assert rs.get_param("effective_variant") == 36 and rs.get_param("rank_mode") == 3
        else:
            assert rs.get_param("effective_variant") == 95 and rs.get_param("rank_mode") == 4
        n = (1 << 19) + 4321
        k = oracle.generate("uniform", n)
        z = oracle.generate("zipf", n)
        v = np.arange(n, dtype=np.uint32)
        for nbits in range(1, 17):
            assert np.array_equal(dev_sort(rs, k, nbits), oracle.sort_keys(k, nbits)), nbits
        for nbits in (8, 7, 6, 5, 4, 11):
            ko, vo = rs.sort_pairs(to_dev(z), to_dev(v), nbits)
            rk, rv = oracle.sort_pairs(z, v, nbits)
            assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv), nbits
        for kind in ("all_equal", "unique16", "sorted", "iota"):
            kk = oracle.generate(kind, n)
            assert np.array_equal(dev_sort(rs, kk, 8), oracle.sort_keys(kk, 8)), kind
        for m in (1, 2, 31, 33, 10367, 10368, 10369, 2 * 10368 + 1,     # around the 10368- and 19456-key tiles
                  19455, 19456, 19457, 2 * 19456 + 1, 9216, 9217):      # (and the 9216-pair tile)
            kk = oracle.generate("uniform", m)
            assert np.array_equal(dev_sort(rs, kk, 8), oracle.sort_keys(kk, 8)), m
        # unaligned input (no bulk copy): same answer
        import torch
        d = to_dev(np.concatenate([np.zeros(1, np.uint32), k]))[1:]
        assert np.array_equal(to_host(rs.sort_keys(d, 8)), oracle.sort_keys(k, 8))
        dv = to_dev(np.concatenate([np.zeros(1, np.uint32), v]))[1:]
        ko, vo = rs.sort_pairs(d, dv, 8)
        rk, rv = oracle.sort_pairs(k, v, 8)
        assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv)
    finally:
        rs.set_param("safe_rank", 0)
        rs.set_param("variant", -1)


@pytest.mark.parametrize("offset_words", [0, 1, 2, 3])
def test_unaligned_arrays_with_a_short_prefetch_distance(rs, oracle, offset_words):
    """The default kernel asks the copy engine to bring a later tile into L2 (16-byte aligned requests only):
    with a distance of one tile every launch of a small sort exercises it; arrays that start 4, 8 or 12 bytes
    off a 16-byte boundary must take the plain path (regression: the host path sorts buckets in place at
    arbitrary offsets)."""
    rs.set_param("prefetch_tiles", 1)
    try:
        n = 5 * 19456 + 777
        k = oracle.generate("uniform", n)
        v = np.arange(n, dtype=np.uint32)
        pad = np.zeros(offset_words, np.uint32)
        d = to_dev(np.concatenate([pad, k]))[offset_words:]
        dv = to_dev(np.concatenate([pad, v]))[offset_words:]
        for nbits in (8, 4):
            assert np.array_equal(to_host(rs.sort_keys(d, nbits)), oracle.sort_keys(k, nbits))
            ko, vo = rs.sort_pairs(d, dv, nbits)
            rk, rv = oracle.sort_pairs(k, v, nbits)
            assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv)
    finally:
        rs.set_param("prefetch_tiles", -1)


@pytest.mark.parametrize("key_bits", [32, 31, 29, 28, 24, 21, 16, 9, 8, 3, 1])
def test_low_bits_sorts_are_full_sorts_when_the_high_bits_agree(rs, oracle, key_bits):
    """b200sort_keys_low_bits / _pairs_low_bits: keys that share their bits >= key_bits (a bucket of an MSD
    partition, a shard of the multi-GPU sort) -- the digits above are skipped, a narrower last digit runs the kernel
    of its own width, and the result is the oracle's full sort."""
    n = 3 * 19456 + 1234
    rng = np.random.default_rng(key_bits)
    low = rng.integers(0, 1 << key_bits, size=n, dtype=np.uint64).astype(np.uint32)
    prefix = np.uint32((0xA5C3F00F >> key_bits) << key_bits) if key_bits < 32 else np.uint32(0)
    k = low | prefix
    v = np.arange(n, dtype=np.uint32)
    for nbits in (8, 5, 11):
        out = rs.sort_keys(to_dev(k), nbits, key_bits=key_bits)
        assert np.array_equal(to_host(out), oracle.sort_keys(k, nbits)), (key_bits, nbits)
    ko, vo = rs.sort_pairs(to_dev(k), to_dev(v), 8, key_bits=key_bits)
    rk, rv = oracle.sort_pairs(k, v, 8)
    assert np.array_equal(to_host(ko), rk) and np.array_equal(to_host(vo), rv)


def test_histogram_matches_tile_table_column_sums(rs, oracle):
    k = oracle.generate("zipf", 300007)
    for shift, bits in ((0, 8), (24, 8), (13, 5), (30, 2), (28, 8)):
        h = rs.histogram(to_dev(k), shift, bits).cpu().numpy().view(np.uint32)
        table, _ = oracle.tile_table(k, 1024, shift, min(bits, 32 - shift))
        expect = np.zeros(1 << bits, np.uint32)
        expect[: table.shape[1]] = table.sum(axis=0)
        assert np.array_equal(h, expect), (shift, bits)


def test_digit_pass_is_one_stable_counting_pass(rs, oracle):
    k = oracle.generate("uniform", 250001)
    v = np.arange(k.size, dtype=np.uint32)
    for shift, bits in ((24, 8), (0, 8), (11, 6), (3, 1)):
        d = (k >> shift) & ((1 << bits) - 1)
        idx = np.argsort(d, kind="stable")
        out = rs.digit_pass(to_dev(k), shift, bits)
        assert np.array_equal(to_host(out), k[idx]), (shift, bits)
        ko, vo = rs.digit_pass(to_dev(k), shift, bits, vals=to_dev(v))
        assert np.array_equal(to_host(ko), k[idx]) and np.array_equal(to_host(vo), v[idx])


def test_digit_pass_with_per_bin_destinations(rs, oracle):
    """bin_dst mode: every bin is written to its own array (what the fused exchange uses)."""
    import torch
    k = oracle.generate("uniform", 123457)
    v = np.arange(k.size, dtype=np.uint32)
    shift, bits = 28, 4
    d = (k >> shift) & 15
    counts = np.bincount(d, minlength=16)
    kbufs = [torch.zeros(int(c) + 1, dtype=torch.int32, device="cuda") for c in counts]
    vbufs = [torch.zeros(int(c) + 1, dtype=torch.int32, device="cuda") for c in counts]
    table = torch.tensor([b.data_ptr() for b in kbufs] + [b.data_ptr() for b in vbufs],
                         dtype=torch.int64, device="cuda")
    rs.digit_pass(to_dev(k), shift, bits, bin_dst=table[:16].contiguous())
    for b in range(16):
        assert np.array_equal(to_host(kbufs[b])[: counts[b]], k[d == b]), b
        kbufs[b].zero_()
    rs.digit_pass(to_dev(k), shift, bits, vals=to_dev(v), bin_dst=table)
    for b in range(16):
        assert np.array_equal(to_host(kbufs[b])[: counts[b]], k[d == b]), b
        assert np.array_equal(to_host(vbufs[b])[: counts[b]], v[d == b]), b
        assert int(kbufs[b][-1]) == 0       # nothing written past the bin


@pytest.mark.parametrize("bulk", [1, 0])
def test_digit_pass_destinations_at_every_alignment(rs, oracle, bulk):
    """Keys with per-bin destinations are written by shared->global bulk copies (16-byte lines) with up to
    three scalar head / tail words per (tile, bin) run: destinations at every 4-byte phase, bins of 1, 2, 3
    and 8 bits (runs from thousands of keys down to a few), guard words around every bin."""
    import torch
    rs.set_param("dst_bulk", bulk)
    try:
        n = 3 * 10368 + 4001
        for kind in ("uniform", "all_equal", "sorted"):
            k = oracle.generate(kind, n)
            for shift, bits in ((31, 1), (30, 2), (29, 3), (24, 8), (0, 8), (5, 3)):
                nb = 1 << bits
                d = (k >> shift) & (nb - 1)
                counts = np.bincount(d, minlength=nb)
                GUARD = 8
                offs, pos = [], 0
                for b in range(nb):
                    pos += GUARD + (b % 4)                 # every 4-byte phase occurs
                    offs.append(pos)
                    pos += int(counts[b])
                arena = torch.full((pos + GUARD,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
                table = torch.tensor([arena.data_ptr() + 4 * o for o in offs], dtype=torch.int64, device="cuda")
                rs.digit_pass(to_dev(k), shift, bits, bin_dst=table)
                got = to_host(arena)
                expect = np.full(pos + GUARD, 0x5A5A5A5A, dtype=np.uint32)
                for b in range(nb):
                    expect[offs[b]: offs[b] + counts[b]] = k[d == b]
                assert np.array_equal(got, expect), (kind, shift, bits, bulk)
    finally:
        rs.set_param("dst_bulk", 1)


def test_exclusive_scan_matches_the_reference_scan_semantics(rs, oracle):
    """b200sort_exclusive_scan == the exclusive scan of the reference's scan() stage
    (SourceCode/Parallel7.cu:485-528; sequential form SourceCode/Baseline1.cu:39-42), mod 2^32."""
    import torch
    for n in (1, 2, 31, 4095, 4096, 4097, 8191, 100003, (1 << 24) + 1):
        x = oracle.generate("uniform", n, first=n) >> 8        # sums wrap modulo 2^32 for the large n
        ref = np.concatenate([[0], np.cumsum(x[:-1], dtype=np.uint64)]).astype(np.uint64) & 0xFFFFFFFF
        got = to_host(rs.exclusive_scan(to_dev(x)))
        assert np.array_equal(got, ref.astype(np.uint32)), n
    # the tile x bin table of a digit pass, bin-major, scanned: Baseline4.cu:127-138
    k = oracle.generate("uniform", 50000)
    table, scan = oracle.tile_table(k, 1024, 8, 4)
    flat = np.ascontiguousarray(table.T).reshape(-1)
    got = to_host(rs.exclusive_scan(to_dev(flat))).reshape(16, -1).T
    assert np.array_equal(got, scan)
    # in place
    d = to_dev(np.ones(70001, np.uint32))
    rs.exclusive_scan(d, out=d)
    assert np.array_equal(to_host(d), np.arange(70001, dtype=np.uint32))
    torch.cuda.synchronize()


def test_device_generators_match_the_oracle(rs, oracle):
    n = 100003
    for kind in oracle.GEN_KINDS:
        dev = rs.generate(kind, n, first=17, total=1 << 20,
                          zipf_cdf=oracle.zipf_cdf() if kind == "zipf" else None)
        assert np.array_equal(to_host(dev), oracle.generate(kind, n, first=17, total=1 << 20)), kind


def test_verify_kernel(rs, oracle):
    k = oracle.generate("uniform", 200001)
    bad, s, h, x = rs.verify(to_dev(k))
    assert bad == int(np.count_nonzero(k[:-1] > k[1:]))
    assert (s, h, x) == oracle.multiset_fingerprint(k)
    bad, s2, h2, x2 = rs.verify(rs.sort_keys(to_dev(k), 8))
    assert bad == 0 and (s2, h2, x2) == (s, h, x)


# ---------------------------------------------------------------------------- full size, by properties

def test_full_size_2_28_properties(rs, oracle):
    """BASELINE config 1 (2^28 uniform keys, 8-bit digits): too big for the oracle to finish in
    seconds, so checked through size-independent properties -- sortedness, the multiset
    fingerprint, idempotence -- plus bit-exact slices against the oracle."""
    import torch
    n = 1 << 28
    d = rs.generate("uniform", n)
    out = rs.sort_keys(d, 8)
    bad, s, h, x = rs.verify(out)
    _, s0, h0, x0 = rs.verify(d)
    assert bad == 0 and (s, h, x) == (s0, h0, x0)
    again = rs.sort_keys(out, 8)
    assert torch.equal(again, out)                       # idempotent
    # uniform keys: the first/last 2^16 outputs are the 2^16 smallest/largest of the input
    host = to_host(d)
    part = np.partition(host, (1 << 16) - 1)[: 1 << 16]
    assert np.array_equal(to_host(out[: 1 << 16]), oracle.sort_keys(part, 8))
    del d, out, again
    torch.cuda.empty_cache()


def test_reference_parallel7_agrees(rs, oracle):
    """The repo's best GPU version (Parallel7 sortByDevice), unmodified, on the same input."""
    if not oracle.ref_available("Parallel7"):
        pytest.skip("oracle/_ref/libref_parallel7.so not present")
    import torch
    n = (1 << 22) + 1
    k = oracle.glibc_rand_keys(n)
    mine = np.zeros_like(k)
    rs.sortByDevice(k, n, mine, 8, 512)
    torch.cuda.synchronize()
    theirs = oracle.ref_sort_by_device(k, 8, 512)
    assert np.array_equal(mine, theirs)
    assert np.array_equal(mine, oracle.sort_keys(k, 8))


def test_device_entry_points_are_reentrant_across_host_threads_and_streams(rs, oracle):
    # include/b200sort.h: the device-pointer entry points may run concurrently given distinct
    # temp buffers and streams (the reference, with its function-static buffers, cannot)
    import threading
    import torch
    rng = np.random.default_rng(21)
    jobs = []
    for t in range(4):
        k = rng.integers(0, 1 << 32, (1 << 20) + 1000 * t + 7, dtype=np.uint64).astype(np.uint32)
        jobs.append({"keys": k, "dev": to_dev(k), "ws": rs.Workspace("cuda"), "stream": torch.cuda.Stream(),
                     "nbits": (8, 4, 8, 5)[t], "out": []})
    torch.cuda.synchronize()

    def work(job):
        with torch.cuda.stream(job["stream"]):
            for _ in range(6):
                job["out"].append(rs.sort_keys(job["dev"], job["nbits"], workspace=job["ws"]))
        job["stream"].synchronize()

    threads = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for j in jobs:
        want = np.sort(j["keys"])
        for o in j["out"]:
            assert np.array_equal(to_host(o), want)


def test_route_counts_thresholds_below_or_equal(rs, oracle):
    import torch
    rng = np.random.default_rng(22)
    k = rng.integers(0, 1 << 32, 100003, dtype=np.uint64).astype(np.uint32)
    k[:5] = [0, 1, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF]
    for thresholds in ([], [0], [1 << 32], [0x80000000], [5, 5, 1 << 31, (1 << 32) - 1, 1 << 32],
                       sorted(int(x) for x in rng.integers(0, 1 << 32, 7, dtype=np.uint64)),
                       sorted(int(x) for x in rng.integers(0, 1 << 32, 200, dtype=np.uint64))):
        got, counts = rs.route(to_dev(k), thresholds, with_counts=True)
        got = to_host(got)
        want = np.searchsorted(np.asarray(thresholds, dtype=np.int64), k.astype(np.int64), side="right")
        assert np.array_equal(got, want.astype(np.uint32)), thresholds[:4]
        assert np.array_equal(to_host(counts), np.bincount(want, minlength=len(thresholds) + 1).astype(np.uint32))
    # tie indices: a key equal to a cut value is at or above the cut from that local index on
    dup = np.repeat(np.array([3, 9, 9, 9, 20], dtype=np.uint32), 1000)
    idx = np.arange(dup.size)
    got = to_host(rs.route(to_dev(dup), [9, 9, 20], [1500, 3200, 1 << 62]))
    want = (dup > 9).astype(np.uint32) * 2 + ((dup == 9) & (idx >= 1500)) + ((dup == 9) & (idx >= 3200))
    assert np.array_equal(got, want.astype(np.uint32))
    # route as the key of a digit pass that carries the real keys: a stable partition by value range
    t = [1 << 30, 1 << 31, 3 << 30]
    r = rs.route(to_dev(k), t)
    _, carried = rs.digit_pass(r, 0, 2, vals=to_dev(k))
    want = k[np.argsort(k >> 30, kind="stable")]
    assert np.array_equal(to_host(carried), want)
