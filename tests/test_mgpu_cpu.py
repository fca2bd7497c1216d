"""CPU: the multi-GPU orchestration (cuda/radixsort_b200/mgpu.py) under world_size-2 gloo.

The device kernels are replaced by numpy stand-ins built on the oracle (test infrastructure), so
what is exercised here is the host logic the product runs unchanged on GPUs: splitter choice,
the G x G count matrix, send/receive split sizes and offsets, stability through the exchange,
and the size-independent verification."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuda.radixsort_b200 import mgpu


def test_choose_owner_uniform_and_skewed():
    hist = np.full(256, 1000)
    for world in (2, 4, 8):
        owner = mgpu.choose_owner(hist, world)
        assert np.all(np.diff(owner) >= 0) and owner[0] == 0 and owner[-1] == world - 1
        assert np.all(np.bincount(owner, minlength=world) == 256 // world)
    skew = np.zeros(256, dtype=np.int64)
    skew[7] = 10 ** 6                      # all keys in one bin: one rank owns everything
    owner = mgpu.choose_owner(skew, 4)
    assert np.all(np.diff(owner) >= 0)
    assert len(set(owner[:8])) <= 2
    assert mgpu.choose_owner(np.zeros(256, dtype=np.int64), 4).shape == (256,)


def test_plan_exchange_is_consistent_across_ranks():
    rng = np.random.default_rng(0)
    world = 4
    counts_all = rng.integers(0, 1000, size=(world, 256))
    plans = [mgpu.plan_exchange(counts_all, r) for r in range(world)]
    for r in range(world):
        p = plans[r]
        assert p["send_counts"].sum() == counts_all[r].sum()
        for s in range(world):
            assert plans[s]["send_counts"][r] == p["recv_counts"][s]          # what s sends is what r expects
        assert p["my_total"] == p["totals"][r] == p["recv_counts"].sum()
    # fused-path offsets tile every owner's buffer exactly, ordered by (source rank, bin)
    owner = plans[0]["owner"]
    for dst in range(world):
        spans = []
        for s in range(world):
            for b in np.nonzero(owner == dst)[0]:
                spans.append((plans[s]["bin_recv_offset"][b], counts_all[s][b], s, b))
        spans.sort()
        pos = 0
        for off, cnt, s, b in spans:
            assert off == pos or cnt == 0
            pos = max(pos, off + cnt)
        assert pos == plans[dst]["my_total"]
        assert [x[2] for x in spans if x[1]] == sorted(x[2] for x in spans if x[1])   # grouped by source rank


def test_shard_key_bits_is_the_common_prefix_of_the_owned_bin_range():
    """mgpu.shard_key_bits: the local sort of a shard may skip the bits on which all its keys agree.  Checked against
    a brute-force common prefix of every key value the shard can hold."""
    rng = np.random.default_rng(7)
    for world in (1, 2, 3, 5, 8):
        for _ in range(20):
            cuts = np.sort(rng.choice(np.arange(1, 256), size=world - 1, replace=False)) if world > 1 else np.array([], int)
            owner = np.searchsorted(cuts, np.arange(256), side="right")
            for shift in (24, 16, 0):
                for r in range(world):
                    kb = mgpu.shard_key_bits(owner, r, shift)
                    mine = np.nonzero(owner == r)[0]
                    lo, hi = int(mine[0]) << shift, ((int(mine[-1]) + 1) << shift) - 1      # key range of the shard
                    assert 1 <= kb <= 32
                    assert (lo >> kb) == (hi >> kb) or kb == 32                               # the promise holds ...
                    if kb > max(1, shift):
                        assert (lo >> (kb - 1)) != (hi >> (kb - 1))                           # ... and is tight
    assert mgpu.shard_key_bits(np.zeros(256, int), 3, 24) == 32                               # a rank that owns nothing


def test_plan_exchange_offsets_match_a_sequential_walk():
    """bin_recv_offset (vectorised per owner) == the obvious walk over the bins in order, for random count
    matrices and for owners that leave some ranks without bins."""
    rng = np.random.default_rng(11)
    for world in (2, 3, 4, 8):
        for trial in range(6):
            bins = 256
            counts = rng.integers(0, 1000, size=(world, bins)).astype(np.int64)
            if trial % 2:                                   # heavy skew: most bins empty, one rank starves
                counts[:, rng.integers(0, bins, 200)] = 0
                counts[:, 7] += 500000
            for rank in range(world):
                plan = mgpu.plan_exchange(counts, rank)
                owner, matrix = plan["owner"], plan["matrix"]
                running = matrix[:rank].sum(axis=0).astype(np.int64) if rank else np.zeros(world, np.int64)
                expect = np.zeros(bins, np.int64)
                for b in range(bins):
                    expect[b] = running[owner[b]]
                    running[owner[b]] += counts[rank, b]
                assert np.array_equal(plan["bin_recv_offset"], expect), (world, trial, rank)
                assert np.array_equal(plan["shard_sizes"], counts.sum(axis=1))


def test_narrow_partition_detection():
    uniform = np.full((4, 256), 100)
    assert mgpu.plan_exchange(uniform, 0)["narrow_bits"] == 2
    assert mgpu.plan_exchange(np.full((2, 256), 7), 1)["narrow_bits"] == 1
    assert mgpu.plan_exchange(np.full((8, 256), 7), 3)["narrow_bits"] == 3
    skew = uniform.copy()
    skew[:, :64] *= 5                       # splitters move off the 2-bit boundaries
    p = mgpu.plan_exchange(skew, 2)
    assert p["narrow_bits"] == 0
    assert mgpu.plan_exchange(np.full((3, 256), 7), 0)["narrow_bits"] == 0    # not a power of two
    # narrow plan offsets: src_base[d] is where this rank's keys start in rank d's buffer
    p = mgpu.plan_exchange(uniform, 2)
    assert np.array_equal(p["src_base"], uniform[:2].reshape(2, 4, 64).sum(axis=2).sum(axis=0))


def _route_np(keys, values, ties):
    """b200sort_route on the host: #{j : (values[j], ties[j]) <= (key, index)}."""
    k = keys.astype(np.int64)
    i = np.arange(k.size, dtype=np.int64)
    dest = np.zeros(k.size, dtype=np.int64)
    for v, c in zip(values, ties):
        dest += (k > int(v)) | ((k == int(v)) & (i >= int(c)))
    return dest


class NumpyOps:
    """numpy stand-ins for the device kernels (same contracts as include/b200sort.h)."""

    def histogram(self, keys, shift, bits):
        k = keys.numpy().view(np.uint32)
        h = np.bincount((k >> shift) & ((1 << bits) - 1), minlength=1 << bits).astype(np.uint32)
        return torch.from_numpy(h.view(np.int32))

    def digit_pass(self, keys, shift, bits, out=None, bin_dst=None, vals=None, out_vals=None):
        k = keys.numpy().view(np.uint32)
        idx = np.argsort((k >> shift) & ((1 << bits) - 1), kind="stable")
        out.numpy().view(np.uint32)[:] = k[idx]
        if vals is None:
            return out
        out_vals.numpy()[:] = vals.numpy()[idx]
        return out, out_vals

    def sort(self, keys, nbits, out, vals=None, out_vals=None, key_bits=32):
        import oracle as O
        k32 = keys.numpy().view(np.uint32)
        if key_bits < 32 and k32.size:      # the promise behind b200sort_keys_low_bits: the bits above agree
            assert np.all((k32 >> np.uint32(key_bits)) == (k32[0] >> np.uint32(key_bits))), key_bits
        if vals is None:
            out.numpy().view(np.uint32)[:] = O.sort_keys(keys.numpy().view(np.uint32), nbits)
            return out
        ko, vo = O.sort_pairs(keys.numpy().view(np.uint32), vals.numpy().view(np.uint32), nbits)
        out.numpy().view(np.uint32)[:] = ko
        out_vals.numpy().view(np.uint32)[:] = vo
        return out, out_vals

    def route(self, keys, values, ties):
        dest = _route_np(keys.numpy().view(np.uint32), values, ties)
        counts = np.bincount(dest, minlength=len(values) + 1)
        return torch.from_numpy(dest.astype(np.int32)), torch.from_numpy(counts.astype(np.int32))

    def sample(self, keys, idx):
        return keys[torch.from_numpy(idx)]

    def empty(self, n):
        return torch.empty(n, dtype=torch.int32)


def _np_verify(t):
    import oracle as O
    k = t.numpy().view(np.uint32)
    s, h, x = O.multiset_fingerprint(k)
    return int(np.count_nonzero(k[:-1] > k[1:])), s, h, x


def _make(O, kind, count, first, total):
    """small16 / small8: uniform keys with constant (zero) high bytes -- the partition digit must move down.
    A "vs:" prefix only changes the sorter (value splitters forced), not the data."""
    kind = kind.split(":")[-1]
    if kind == "heavyvalue":    # 70 % of the keys are ONE value: its run must be split between the ranks
        k = O.generate("uniform", count, first=first, total=total)
        heavy = (O.generate("uniform", count, first=first + total, total=2 * total) % np.uint32(10)) < 7
        return np.where(heavy, np.uint32(0x5A5A5A5A), k)
    if kind == "heavybin":      # 90 % of the keys share one top byte but differ below it
        k = O.generate("uniform", count, first=first, total=total)
        heavy = (O.generate("uniform", count, first=first + total, total=2 * total) % np.uint32(10)) != 0
        return np.where(heavy, (k & np.uint32(0x00FFFFFF)) | np.uint32(0x5A000000), k)
    if kind == "small16":
        return O.generate("uniform", count, first=first, total=total) & np.uint32(0xFFFF)
    if kind == "small8":
        return O.generate("uniform", count, first=first, total=total) & np.uint32(0xFF)
    return O.generate(kind, count, first=first, total=total)


def _worker(rank, world, port, kind, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    per = n_total // world
    first = rank * per
    count = per if rank < world - 1 else n_total - first
    keys = torch.from_numpy(_make(O, kind, count, first, n_total).view(np.int32))
    sorter = mgpu.ShardedSorter(dist.group.WORLD, nbits=8, ops=NumpyOps(), time_phases=False,
                                balance_threshold=1.0 if kind.startswith("vs:") else 1.2)
    res = sorter.sort(keys)
    key_shift = sorter.partition_shift
    by_value = "value_thresholds" in sorter.last_plan
    ok = mgpu.verify_sharded(res, keys, verify_fn=_np_verify)
    # a corrupted shard must be caught
    if res.numel() > 2:
        broken = res.clone()
        if int(res[0]) != int(res[-1]):
            broken[0], broken[-1] = res[-1], res[0]        # order violation
        else:
            broken[0] = int(res[0]) ^ 1                     # constant shard: change the multiset instead
        caught = not mgpu.verify_sharded(broken, keys, verify_fn=_np_verify)
    else:
        caught = not mgpu.verify_sharded(torch.cat([res, res.new_zeros(1)]), keys, verify_fn=_np_verify)
    np.save(os.path.join(out_dir, f"shard{rank}.npy"), res.numpy().view(np.uint32))   # before buffers are reused
    # key/value variant: value = global input index, keys with many duplicates
    kk = torch.from_numpy((keys.numpy().view(np.uint32) & np.uint32(0xFF0000FF)).view(np.int32).copy())
    vv = torch.arange(first, first + count, dtype=torch.int32)
    pk, pv = sorter.sort_pairs(kk, vv)
    np.save(os.path.join(out_dir, f"pairs_k{rank}.npy"), pk.numpy().view(np.uint32).copy())
    np.save(os.path.join(out_dir, f"pairs_v{rank}.npy"), pv.numpy().view(np.uint32).copy())
    np.save(os.path.join(out_dir, f"flags{rank}.npy"), np.array([ok, caught, key_shift, by_value,
                                                                 "value_thresholds" in sorter.last_plan]))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("kind,n_total", [("uniform", 200003), ("zipf", 120001), ("all_equal", 5000),
                                          ("sorted", 70000), ("small16", 90001), ("small8", 30000),
                                          ("heavybin", 150001), ("vs:uniform", 100003), ("vs:zipf", 90001),
                                          ("vs:unique16", 60000), ("vs:sorted", 50001), ("heavyvalue", 80001)])
def test_sharded_sort_world2_gloo(tmp_path, kind, n_total):
    import oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), kind, n_total, str(tmp_path)), nprocs=world, join=True)
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    flags = [np.load(tmp_path / f"flags{r}.npy") for r in range(world)]
    whole = _make(O, kind, n_total, 0, n_total)
    assert np.array_equal(np.concatenate(shards), O.sort_keys(whole, 8))       # concatenation in rank order
    assert all(f[0] for f in flags), "verify_sharded rejected a correct result"
    assert all(f[1] for f in flags), "verify_sharded accepted a corrupted result"
    if kind in ("uniform", "small16", "small8", "all_equal", "heavybin", "vs:uniform", "vs:sorted"):
        assert abs(len(shards[0]) - len(shards[1])) < 0.03 * n_total + 2          # balanced split
    if kind == "heavyvalue":   # the run of the heavy value is cut at a position: not the 10.6 % / 89.4 % of an unsplit run
        assert abs(len(shards[0]) - len(shards[1])) < 0.03 * n_total
    expected_shift = {"small16": 8, "small8": 0, "all_equal": 0}.get(kind, 24)
    assert all(int(f[2]) == expected_shift for f in flags)                        # partition digit moved down
    if kind.startswith("vs:") or kind in ("heavybin", "heavyvalue"):
        assert all(int(f[3]) for f in flags), "value splitters were expected for the keys sort"
    # pairs: concatenation == stable sort of (masked key, global index)
    pk = np.concatenate([np.load(tmp_path / f"pairs_k{r}.npy") for r in range(world)])
    pv = np.concatenate([np.load(tmp_path / f"pairs_v{r}.npy") for r in range(world)])
    mk = whole & np.uint32(0xFF0000FF)
    rk, rv = O.sort_pairs(mk, np.arange(n_total, dtype=np.uint32), 8)
    assert np.array_equal(pk, rk) and np.array_equal(pv, rv)


@pytest.mark.parametrize("world,kind,n_total", [(4, "uniform", 160003), (4, "vs:zipf", 100001), (4, "heavyvalue", 120001),
                                                (3, "small16", 90002), (3, "heavybin", 90001)])
def test_sharded_sort_more_ranks_gloo(tmp_path, world, kind, n_total):
    """Same check with 3 ranks (no narrow partition: not a power of two) and 4 ranks."""
    import oracle as O
    mp.spawn(_worker, args=(world, _free_port(), kind, n_total, str(tmp_path)), nprocs=world, join=True)
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    flags = [np.load(tmp_path / f"flags{r}.npy") for r in range(world)]
    whole = _make(O, kind, n_total, 0, n_total)
    assert np.array_equal(np.concatenate(shards), O.sort_keys(whole, 8))
    assert all(f[0] for f in flags) and all(f[1] for f in flags)
    sizes = np.array([len(x) for x in shards])
    assert sizes.max() < 1.12 * n_total / world + 64, sizes                     # every kind here must end up balanced
    if kind != "uniform" and kind != "small16":
        assert all(int(f[3]) for f in flags), "value splitters were expected"
    pk = np.concatenate([np.load(tmp_path / f"pairs_k{r}.npy") for r in range(world)])
    pv = np.concatenate([np.load(tmp_path / f"pairs_v{r}.npy") for r in range(world)])
    rk, rv = O.sort_pairs(whole & np.uint32(0xFF0000FF), np.arange(n_total, dtype=np.uint32), 8)
    assert np.array_equal(pk, rk) and np.array_equal(pv, rv)


def _route_all(shards, world, m=2048):
    """Host model of the value-splitter plan: returns per-shard destination arrays."""
    positions = [mgpu.sample_indices(sh.size, m) for sh in shards]
    samples = [sh[p] if sh.size else np.zeros(0, np.uint32) for sh, p in zip(shards, positions)]
    pool = np.sort(np.concatenate(samples))
    values, split_rank, split_pos = mgpu.value_splitters(pool, samples, positions, world)
    dests = []
    for r, sh in enumerate(shards):
        v, c = mgpu.thresholds_for_rank(values, split_rank, split_pos, r)
        assert np.all(np.diff(v) >= 0)
        dests.append(_route_np(sh, v, c))
    return dests


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_value_splitters_give_a_stable_balanced_range_partition(world):
    rng = np.random.default_rng(5)
    n = 40000
    cases = {
        "uniform": rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32),
        "all_equal": np.full(n, 0xDEADBEEF, dtype=np.uint32),
        "max_value": np.full(n, 0xFFFFFFFF, dtype=np.uint32),
        "two_values": rng.integers(0, 2, n).astype(np.uint32) * np.uint32(77),
        "heavy_head": np.where(rng.random(n) < 0.6, np.uint32(123456), rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)),
        "sixteen": (rng.integers(0, 16, n).astype(np.uint32) * np.uint32(0x01010101)),
    }
    for name, keys in cases.items():
        shards = np.array_split(keys, world)
        dests = _route_all(shards, world)
        dest = np.concatenate(dests)
        # a range partition: destinations are monotone in the key ...
        order = np.argsort(keys, kind="stable")
        assert np.all(np.diff(dest[order]) >= 0), name
        # ... which for equal keys means monotone in the global index (the stable order)
        shares = np.bincount(dest, minlength=world)
        assert shares.max() <= 1.10 * n / world + 64, (name, world, shares)      # sampling noise only


def test_sample_indices():
    for n_local in (1, 7, 8192, 8193, 10 ** 6 + 3):
        idx = mgpu.sample_indices(n_local, 8192)
        assert idx.size == 8192 and idx.min() >= 0 and idx.max() < n_local and np.all(np.diff(idx) >= 0)
    assert mgpu.sample_indices(0).size == 0
    v, r, c = mgpu.value_splitters(np.zeros(0, np.uint32), [np.zeros(0, np.uint32)] * 4, [np.zeros(0, np.int64)] * 4, 4)
    assert v.shape == (3,) and r.shape == (3,) and c.shape == (3,)


def test_c_planners_agree_with_the_python_driver(rs):
    """The single-process driver (csrc/mgpu_host.cu) and the torch.distributed driver (mgpu.py) plan
    with separate implementations of the same rules; both are exported/pure, so compare them here."""
    import ctypes as C
    lib = rs.load()
    rng = np.random.default_rng(31)
    for trial in range(40):
        world = int(rng.choice([2, 4, 8, 16]))
        hist = rng.integers(0, 1000, 256).astype(np.uint64)
        if trial % 3 == 0:
            hist[rng.integers(0, 256)] += np.uint64(rng.integers(10 ** 4, 10 ** 6))    # a heavy bin
        if trial % 7 == 0:
            hist[:] = 0
            hist[rng.integers(0, 256)] = 12345                                           # a single bin
        owner = np.zeros(256, dtype=np.int32)
        assert lib.b200sort_plan_owners(hist.ctypes.data, 256, world, owner.ctypes.data) == 0
        assert np.array_equal(owner, mgpu.choose_owner(hist.astype(np.int64), world)), (trial, world)
    for trial in range(40):
        world = int(rng.choice([2, 3, 4, 8]))
        m = int(rng.choice([64, 257, 1024]))
        sizes = [int(rng.integers(m, 50 * m)) for _ in range(world)]
        if trial % 5 == 0:
            sizes[int(rng.integers(0, world))] = 0                                       # an empty shard
        kind = trial % 4
        samples, positions = [], []
        for n_local in sizes:
            pos = mgpu.sample_indices(n_local, m)
            if kind == 0:
                keys = rng.integers(0, 1 << 32, pos.size, dtype=np.uint64)
            elif kind == 1:
                keys = rng.integers(0, 5, pos.size, dtype=np.uint64) * 1000                # few values
            elif kind == 2:
                keys = np.where(rng.random(pos.size) < 0.7, 0xFFFFFFFF, rng.integers(0, 1 << 32, pos.size, dtype=np.uint64))
            else:
                keys = np.full(pos.size, 42, dtype=np.uint64)                               # all equal
            samples.append(keys.astype(np.uint32))
            positions.append(pos.astype(np.uint64))
        pool = np.sort(np.concatenate(samples))
        want_v, want_r, want_p = mgpu.value_splitters(pool, samples, positions, world)
        cat_k = np.ascontiguousarray(np.concatenate(samples))
        cat_p = np.ascontiguousarray(np.concatenate(positions))
        offs = np.concatenate([[0], np.cumsum([s.size for s in samples])]).astype(np.uint64)
        v = np.zeros(world - 1, np.uint64); r = np.zeros(world - 1, np.int32); p = np.zeros(world - 1, np.uint64)
        assert lib.b200sort_plan_value_cuts(cat_k.ctypes.data, cat_p.ctypes.data, offs.ctypes.data, world,
                                            v.ctypes.data, r.ctypes.data, p.ctypes.data) == 0
        assert np.array_equal(v.astype(np.int64), want_v), (trial, kind, world)
        assert np.array_equal(r.astype(np.int64), want_r), (trial, kind, world)
        assert np.array_equal(p.astype(np.int64), want_p), (trial, kind, world)
    assert lib.b200sort_plan_owners(None, 256, 2, None) == -1
