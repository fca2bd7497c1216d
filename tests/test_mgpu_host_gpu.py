"""GPU parity of the single-process multi-GPU host entry points (b200sort_mgpu_*_host) against
the oracle.  Shards may share one GPU (a repeated ordinal), so the whole partition / exchange /
local-sort path runs on a one-GPU box as well; with two or more GPUs the same cases also run
across real peers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def device_sets():
    import torch
    sets = [[0], [0, 0], [0, 0, 0], [0, 0, 0, 0]]
    g = torch.cuda.device_count()
    if g >= 2:
        sets.append(list(range(2)))
    if g >= 4:
        sets.append(list(range(4)))
    if g >= 3:
        sets.append(list(range(g)))
    return sets


def run(rs, keys, nbits, devices):
    out = np.zeros_like(keys)
    rs.sort_by_devices(keys, keys.size, out, nbits, 512, devices)
    return out


def test_kat_debug_and_default_config(rs, oracle):
    k = oracle.glibc_rand_keys(513, 0xFF)
    for devs in device_sets():
        assert oracle.fnv1a64(run(rs, k, 4, devs)) == 0x714658018BFDCBDC
    k = oracle.glibc_rand_keys((1 << 24) + 1)
    for devs in device_sets()[1:]:
        out = run(rs, k, 8, devs)
        assert oracle.fnv1a64(out) == 0xE354BCFF33580302
        st = rs.mgpu_last_stats()
        assert st["devices"] == len(devs) and st["imbalance"] < 1.1


@pytest.mark.parametrize("kind", ["uniform", "zipf", "unique16", "all_equal", "sorted", "reversed", "iota"])
def test_distributions(rs, oracle, kind):
    n = (1 << 20) + 77
    k = oracle.generate(kind, n)
    want = oracle.sort_keys(k, 8)
    for devs in device_sets():
        assert np.array_equal(run(rs, k, 8, devs), want), (kind, devs)


def test_uniform_keys_take_the_narrow_partition(rs, oracle):
    k = oracle.generate("uniform", 1 << 21)
    out = run(rs, k, 8, [0, 0, 0, 0])
    assert np.array_equal(out, oracle.sort_keys(k, 8))
    st = rs.mgpu_last_stats()
    assert st["partition_bits"] == 2 and st["partition_shift"] == 30 and st["imbalance"] < 1.05
    run(rs, k, 8, [0, 0, 0])
    assert rs.mgpu_last_stats()["partition_bits"] == 8


def test_edge_sizes_and_digit_widths(rs, oracle):
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 3, 5, 31, 32, 33, 1000, 11264, 11265, (1 << 18) + 1):
        k = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        for devs in ([0, 0], [0, 0, 0]):
            assert np.array_equal(run(rs, k, 8, devs), np.sort(k)), (n, devs)
    k = rng.integers(0, 1 << 32, 300001, dtype=np.uint64).astype(np.uint32)
    for nbits in (1, 3, 4, 5, 7, 11, 16):
        assert np.array_equal(run(rs, k, nbits, [0, 0]), oracle.sort_keys(k, nbits)), nbits


def test_keys_sharing_their_high_bytes(rs, oracle):
    rng = np.random.default_rng(12)
    n = 200003
    for low_bits in (24, 16, 8, 3):
        k = (np.uint32(0xA5C3F00F) & np.uint32((0xFFFFFFFF << low_bits) & 0xFFFFFFFF)) | \
            rng.integers(0, 1 << low_bits, n, dtype=np.uint64).astype(np.uint32)
        out = run(rs, k, 8, [0, 0, 0, 0])
        assert np.array_equal(out, np.sort(k)), low_bits
        assert rs.mgpu_last_stats()["partition_shift"] < 32 - 8 or low_bits > 24


def test_pairs_are_stable_across_shards(rs, oracle):
    rng = np.random.default_rng(13)
    n = (1 << 19) + 5
    for hi in (1 << 32, 1 << 10, 3):
        k = rng.integers(0, hi, n, dtype=np.uint64).astype(np.uint32)
        v = np.arange(n, dtype=np.uint32)
        wk, wv = oracle.sort_pairs(k, v, 8)
        for devs in device_sets():
            ok, ov = np.zeros_like(k), np.zeros_like(v)
            rs.sort_pairs_by_devices(k, v, n, ok, ov, 8, 512, devs)
            assert np.array_equal(ok, wk) and np.array_equal(ov, wv), (hi, devs)


def test_pageable_and_pinned_hosts_and_in_place(rs, oracle):
    import torch
    n = (1 << 24) + 4099          # several 16 MiB staging chunks per shard, ragged tail
    k = oracle.generate("uniform", n)
    want = np.sort(k)
    for devs in ([0, 0], [0, 0, 0]):
        assert np.array_equal(run(rs, k, 8, devs), want)
    pin_in = torch.empty(n, dtype=torch.int32).pin_memory()
    pin_out = torch.empty(n, dtype=torch.int32).pin_memory()
    pin_in.numpy().view(np.uint32)[:] = k
    rs.sort_by_devices(pin_in.numpy().view(np.uint32), n, pin_out.numpy().view(np.uint32), 8, 512, [0, 0])
    assert np.array_equal(pin_out.numpy().view(np.uint32), want)
    buf = k.copy()
    rs.sort_by_devices(buf, n, buf, 8, 512, [0, 0])   # h_out == h_in, as the reference tolerates
    assert np.array_equal(buf, want)


def test_bad_devices_are_refused(rs):
    k = np.arange(8, dtype=np.uint32)
    with pytest.raises(rs.RadixSortError):
        rs.sort_by_devices(k, 8, k.copy(), 8, 512, [0, 4096])
    rs.sort_by_devices(k, 8, k.copy(), 8, 512, None)   # every visible device


def heavy_bin_keys(oracle, n):
    """90 % of the keys share one top byte and differ below it: no edge of the 256 top-byte bins
    can balance the shards."""
    k = oracle.generate("uniform", n)
    heavy = (oracle.generate("uniform", n, first=n, total=2 * n) % np.uint32(10)) != 0
    return np.where(heavy, (k & np.uint32(0x00FFFFFF)) | np.uint32(0x5A000000), k).astype(np.uint32)


def test_skewed_keys_switch_to_value_splitters(rs, oracle):
    n = (1 << 21) + 11
    k = heavy_bin_keys(oracle, n)
    want = np.sort(k)
    for devs in device_sets()[1:]:
        assert np.array_equal(run(rs, k, 8, devs), want), devs
        st = rs.mgpu_last_stats()
        assert st["value_splitters"] == 1 and st["imbalance"] < 1.1, (devs, st)
    # the same input with the switch disabled: still correct, but one shard takes most of the keys
    rs.set_param("mgpu_balance_permille", 0)
    try:
        assert np.array_equal(run(rs, k, 8, [0, 0, 0, 0]), want)
        st = rs.mgpu_last_stats()
        assert st["value_splitters"] == 0 and st["imbalance"] > 2.0
    finally:
        rs.set_param("mgpu_balance_permille", 1200)
    # forced on for balanced and duplicate-heavy inputs alike
    rs.set_param("mgpu_balance_permille", 1)
    try:
        for kind in ("uniform", "zipf", "unique16", "all_equal", "sorted", "reversed"):
            kk = oracle.generate(kind, (1 << 19) + 3)
            for devs in ([0, 0], [0, 0, 0], [0, 0, 0, 0, 0]):
                assert np.array_equal(run(rs, kk, 8, devs), np.sort(kk)), (kind, devs)
                assert rs.mgpu_last_stats()["value_splitters"] == 1
    finally:
        rs.set_param("mgpu_balance_permille", 1200)


def test_value_splitters_keep_pairs_stable(rs, oracle):
    n = (1 << 20) + 5
    k = heavy_bin_keys(oracle, n) & np.uint32(0xFFFF00FF)      # plenty of duplicates inside the heavy bin
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v, 8)
    for devs in device_sets()[1:]:
        ok, ov = np.zeros_like(k), np.zeros_like(v)
        rs.sort_pairs_by_devices(k, v, n, ok, ov, 8, 512, devs)
        assert rs.mgpu_last_stats()["value_splitters"] == 1
        assert np.array_equal(ok, wk) and np.array_equal(ov, wv), devs


def test_one_heavy_value_is_split_between_shards_by_source_order(rs, oracle):
    n = (1 << 21) + 3
    k = oracle.generate("uniform", n)
    heavy = (oracle.generate("uniform", n, first=n, total=2 * n) % np.uint32(10)) < 7
    k = np.where(heavy, np.uint32(0x5A5A5A5A), k).astype(np.uint32)          # 70 % of the keys are one value
    v = np.arange(n, dtype=np.uint32)
    wk, wv = oracle.sort_pairs(k, v, 8)
    for devs in ([0, 0, 0, 0], [0] * 8) + tuple(d for d in device_sets() if len(set(d)) > 1):
        ok, ov = np.zeros_like(k), np.zeros_like(v)
        rs.sort_pairs_by_devices(k, v, n, ok, ov, 8, 512, list(devs))
        st = rs.mgpu_last_stats()
        assert np.array_equal(ok, wk) and np.array_equal(ov, wv), devs      # ties still in input order
        assert st["value_splitters"] == 1 and st["imbalance"] < 1.06, (devs, st)   # the run is cut at a position
    # all keys equal: every shard keeps its own keys
    kk = np.full(n, 7, dtype=np.uint32)
    assert np.array_equal(run(rs, kk, 8, [0, 0, 0, 0]), kk)
    assert rs.mgpu_last_stats()["imbalance"] < 1.01


def test_small_skewed_pairs_after_large_keys_only_call(rs, oracle):
    """Regression (round-1 advisor finding): the local sort may reuse the shard INPUT buffers as its
    output; a large keys-only call grows only the key input buffer, so a later, smaller pairs call whose
    received range exceeds its own shard (skewed keys) must not write values past a small value buffer."""
    rs.mgpu_shutdown()                                   # start from empty per-shard buffers
    devs = [0, 0, 0, 0]
    big = oracle.generate("uniform", 1 << 22)
    assert np.array_equal(run(rs, big, 8, devs), oracle.sort_keys(big, 8))
    # tiny pairs call first so that the value buffers exist but are small
    k0 = oracle.generate("uniform", 1 << 10)
    v0 = np.arange(k0.size, dtype=np.uint32)
    ko = np.zeros_like(k0); vo = np.zeros_like(v0)
    rs.sort_pairs_by_devices(k0, v0, k0.size, ko, vo, 8, 512, devs)
    # skewed pairs: the top-byte plan gives one shard far more than n/4 pairs
    n = 1 << 18
    k = oracle.generate("uniform", n)
    k[: n - n // 8] = (k[: n - n // 8] & 0x00FFFFFF) | 0x7F000000
    v = np.arange(n, dtype=np.uint32)
    ko = np.zeros_like(k); vo = np.zeros_like(v)
    rs.set_param("mgpu_balance_permille", 0)             # keep the bin-edge plan: one shard receives 7/8 of the pairs
    try:
        rs.sort_pairs_by_devices(k, v, n, ko, vo, 8, 512, devs)
    finally:
        rs.set_param("mgpu_balance_permille", 1200)
    rk, rv = oracle.sort_pairs(k, v, 8)
    assert np.array_equal(ko, rk) and np.array_equal(vo, rv)
    assert rs.mgpu_last_stats()["imbalance"] > 2.0
