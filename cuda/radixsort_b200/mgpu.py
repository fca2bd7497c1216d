"""Sharded sort over 2/4/8 GPUs: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch) for the plumbing, libb200sort kernels for every byte of compute.

The reference has nothing multi-GPU; the parity target is
"concatenation of the shards in rank order == sortByHost of the whole input".

    top-digit histogram (b200sort_histogram, 8 bits)            local,  4 B/key read
    all_gather of the G x 256 count matrix                      16 KiB, latency bound
    splitters = top-digit bin boundaries nearest to j*N/G       host, identical on every rank
    MSD partition = ONE stable digit pass on the top digit      local,  8 B/key
        - exchange=nccl : into a local buffer, then one all_to_all_single over NVLink
        - exchange=fused: the digit-pass kernel stores every bin straight into the owning
          rank's receive buffer (symmetric memory, peer pointers) -- partition and exchange
          are one kernel, the NVLink transfer overlaps the ranking tile by tile
    local LSD sort of the received bucket (b200sort_keys)        local,  36 B/key

Receive offsets are ordered by source rank and every step is stable, so equal keys keep their
global input order (matters for the key/value variant).

Skewed keys (bin edges would leave a shard above balance_threshold x the mean): the splitters
become cuts of the key space taken from a sample -- value_splitters() below -- and the
partition runs on a route array (b200sort_route: destination of every key) that carries the
real keys; see ShardedSorter._sort_by_value_splitters.  The single-process twin of this driver
is csrc/mgpu_host.cu (b200sort_mgpu_*_host); tests compare the two planners.

`ops` abstracts the device kernels so that the orchestration (everything in this file) can be
exercised on CPU with the gloo backend in tests (tests/test_mgpu_cpu.py supplies numpy ops);
the product always uses DeviceOps.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

TOP_BITS = 8


# ---------------------------------------------------------------------------------------------
# pure host logic (unit-tested on CPU)

def choose_owner(global_hist: np.ndarray, world: int) -> np.ndarray:
    """owner[b] = rank that receives top-digit bin b.  Boundaries are the bin edges closest to
    the ideal cut points j*N/world, so owners are non-decreasing in b and every rank gets a
    contiguous range of bins (possibly empty for heavily skewed inputs)."""
    bins = global_hist.size
    total = int(global_hist.sum())
    csum = np.concatenate([[0], np.cumsum(global_hist.astype(np.int64))])  # csum[b] = keys in bins < b
    cuts = [0]
    for j in range(1, world):
        target = j * total / world
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, 0), bins)
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[b] - target):
            b -= 1
        cuts.append(max(b, cuts[-1]))
    cuts.append(bins)
    owner = np.empty(bins, dtype=np.int64)
    for r in range(world):
        owner[cuts[r]:cuts[r + 1]] = r
    return owner


def shard_key_bits(owner: np.ndarray, rank: int, shift: int) -> int:
    """Bits of a key that can still differ inside rank's shard after a range partition on the TOP_BITS-wide window
    at `shift` (all keys agree above the window, or the window would sit higher): the bins a rank owns are a range
    [lo, hi] of window values, which agree in their leading bits above the highest bit of lo ^ hi."""
    mine = np.nonzero(np.asarray(owner) == rank)[0]
    if mine.size == 0:
        return 32
    lo, hi = int(mine[0]), int(mine[-1])
    return min(32, shift + int(lo ^ hi).bit_length()) if lo != hi else max(1, shift)


def plan_exchange(counts_all: np.ndarray, rank: int, owner=None):
    """counts_all[src][bin] -> everything a rank needs for the exchange.

    Returns dict with owner[bin], send_counts[dst], recv_counts[src], recv_offsets[src] (where
    src's keys start in my receive buffer), my_total, and bin_recv_offset[bin]: the offset, in
    the OWNER's receive buffer, where this rank's keys of `bin` start (used by the fused path)."""
    counts_all = np.asarray(counts_all, dtype=np.int64)
    world, bins = counts_all.shape
    owner = choose_owner(counts_all.sum(axis=0), world) if owner is None else np.asarray(owner, dtype=np.int64)
    matrix = np.zeros((world, world), dtype=np.int64)  # matrix[src][dst]
    for dst in range(world):
        matrix[:, dst] = counts_all[:, owner == dst].sum(axis=1)
    send_counts = matrix[rank].copy()
    recv_counts = matrix[:, rank].copy()
    recv_offsets = np.concatenate([[0], np.cumsum(recv_counts)[:-1]])
    # offset of (rank, bin) inside owner's buffer = keys from lower source ranks for that owner
    #                                             + this rank's keys of earlier bins of the same owner
    src_base = np.zeros(world, dtype=np.int64)  # src_base[dst] = sum_{s < rank} matrix[s][dst]
    if rank > 0:
        src_base = matrix[:rank].sum(axis=0)
    bin_recv_offset = np.zeros(bins, dtype=np.int64)
    mine = counts_all[rank]
    for d in range(world):              # vectorised per owner (the planner runs while the GPU waits)
        sel = owner == d
        c = mine[sel]
        bin_recv_offset[sel] = src_base[d] + np.cumsum(c) - c
    # Splitters that coincide with the top log2(world) bits (uniform keys on 2/4/8 ranks): the
    # partition can then run as a log2(world)-bit digit pass -- 2..8 bins instead of 256, i.e.
    # runs of thousands of keys per (tile, destination), which is what NVLink stores want.
    lg = int(world).bit_length() - 1
    narrow_bits = 0
    if world > 1 and (1 << lg) == world and bins >= world:
        shift = int(bins).bit_length() - 1 - lg
        if np.array_equal(owner, np.arange(bins) >> shift):
            narrow_bits = lg
    return {"owner": owner, "matrix": matrix, "send_counts": send_counts, "recv_counts": recv_counts,
            "recv_offsets": recv_offsets, "my_total": int(recv_counts.sum()),
            "bin_recv_offset": bin_recv_offset, "totals": matrix.sum(axis=0), "shard_sizes": matrix.sum(axis=1),
            "src_base": src_base, "narrow_bits": narrow_bits}


SAMPLE_PER_RANK = 8192


_sample_cache: dict = {}


def sample_indices(n_local: int, m: int = SAMPLE_PER_RANK) -> np.ndarray:
    """m positions spread over a shard: one per stride, at a position inside the stride that
    changes from sample to sample (so periodic inputs do not alias with the stride)."""
    if n_local <= 0:
        return np.zeros(0, dtype=np.int64)
    hit = _sample_cache.get((n_local, m))
    if hit is not None:
        return hit
    i = np.arange(m, dtype=np.int64)
    base = (i * n_local) // m
    stride = max(n_local // m, 1)
    jitter = ((i * 2654435761) & 0xFFFFFFFF) % stride
    out = np.minimum(base + jitter, n_local - 1)
    if len(_sample_cache) < 64:
        _sample_cache[(n_local, m)] = out
    return out


TIE_ALL_LEFT = 1 << 62   # tie index beyond any shard: every key equal to the cut value stays left of the cut


def value_splitters(sorted_pool: np.ndarray, samples_by_rank, positions_by_rank, world: int):
    """Cuts of the key space taken from a sample: returns (values, split_rank, split_pos), each of
    world-1 entries.

    Cut j aims at the j/world quantile of the pooled sample and sits at a key VALUE v_j.  Keys
    below v_j go left of the cut, keys above go right.  Keys EQUAL to v_j go left when they come
    from a source rank < split_rank[j], right from a rank > split_rank[j], and inside rank
    split_rank[j] those at local index < split_pos[j] go left.  Cutting a run of equal keys at a
    position of the global input order keeps ties in that order (lower ranks hold lower global
    indices), so the sort stays stable while one heavy value -- or an all-equal input -- spreads
    over as many shards as its size asks for.  positions_by_rank[r][i] is the local index the
    i-th sample of rank r was taken from (ascending)."""
    s = np.asarray(sorted_pool).astype(np.int64)
    m = s.size
    values = np.zeros(max(world - 1, 0), dtype=np.int64)
    split_rank = np.zeros(max(world - 1, 0), dtype=np.int64)
    split_pos = np.zeros(max(world - 1, 0), dtype=np.int64)
    if m == 0:
        return values, split_rank, split_pos
    per_rank = [np.asarray(x) for x in samples_by_rank]
    for j in range(1, world):
        q = min((j * m) // world, m - 1)
        v = int(s[q])
        lo = int(np.searchsorted(s, v, side="left"))           # lo <= q < hi: the cut lies inside the run of v
        left = q - lo                                           # sampled copies of v that belong left of the cut
        values[j - 1] = v
        split_rank[j - 1] = world                                # default: the whole run goes left
        if left == 0:                                            # the cut sits at the start of the run: all of it goes right
            split_rank[j - 1] = 0
            continue
        for r in range(world):
            hits = np.flatnonzero(per_rank[r] == v)
            if left < hits.size:
                split_rank[j - 1] = r
                split_pos[j - 1] = int(positions_by_rank[r][hits[left]])
                break
            left -= hits.size
    return values, split_rank, split_pos


def thresholds_for_rank(values, split_rank, split_pos, rank: int):
    """(values, ties) of one source rank for b200sort_route: a key equal to values[j] is at or above
    cut j from local index ties[j] on."""
    split_rank = np.asarray(split_rank, dtype=np.int64)
    ties = np.where(rank < split_rank, TIE_ALL_LEFT, np.where(rank > split_rank, 0, np.asarray(split_pos, dtype=np.int64)))
    return np.asarray(values, dtype=np.int64), ties.astype(np.int64)


# ---------------------------------------------------------------------------------------------
class DeviceOps:
    """The product's device operations: libb200sort through the C ABI."""

    def __init__(self):
        from . import api
        self.api = api
        self.ws = api.Workspace("cuda")

    def histogram(self, keys, shift, bits):
        return self.api.histogram(keys, shift, bits, workspace=self.ws)

    def digit_pass(self, keys, shift, bits, out=None, bin_dst=None, vals=None, out_vals=None):
        return self.api.digit_pass(keys, shift, bits, vals=vals, out_keys=out, out_vals=out_vals,
                                   bin_dst=bin_dst, workspace=self.ws)

    def sort(self, keys, nbits, out, vals=None, out_vals=None, key_bits=32):
        if vals is None:
            return self.api.sort_keys(keys, nbits, out=out, workspace=self.ws, key_bits=key_bits)
        return self.api.sort_pairs(keys, vals, nbits, out_keys=out, out_vals=out_vals, workspace=self.ws,
                                   key_bits=key_bits)

    def route(self, keys, values, ties):
        return self.api.route(keys, values, ties, with_counts=True)

    def sample(self, keys, idx):
        key = (keys.numel(), idx.size)
        if getattr(self, "_sample_key", None) != key:       # positions depend on the shard size only
            self._sample_key, self._sample_idx = key, torch.from_numpy(idx).to(keys.device)
        return keys[self._sample_idx]

    def empty(self, n):
        return torch.empty(n, dtype=torch.int32, device="cuda")


class PhaseTimer:
    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.marks = []

    def mark(self, name):
        if self.enabled:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def report(self):
        if not self.enabled or len(self.marks) < 2:
            return {}
        torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


class ShardedSorter:
    """Sorts a uint32 array (optionally with uint32 values, stably) that is sharded over the ranks
    of `group`."""

    def __init__(self, group=None, per_rank_capacity: int = 0, nbits: int = 8, fused: bool = False,
                 ops=None, time_phases: bool = True, allow_narrow: bool = True, balance_threshold: float = 1.2):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.nbits = nbits
        self.ops = ops if ops is not None else DeviceOps()
        self.on_gpu = ops is None
        self.capacity = int(per_rank_capacity)
        self.fused = bool(fused) and self.on_gpu
        self.allow_narrow = allow_narrow
        # bin-edge splitters leaving one shard above balance_threshold x the mean switch the sort
        # to value splitters taken from a sample (0 = never)
        self.balance_threshold = float(balance_threshold)
        self.route_dump = None
        self.timer = PhaseTimer(time_phases and self.on_gpu)
        self.last_plan = None
        self.recv = None
        self.peer_ptrs = None
        self.symm = None
        self.part = None
        self.out = None
        self.recv_v = None
        self.peer_ptrs_v = None
        self.symm_v = None
        self.part_v = None
        self.out_v = None
        if self.capacity:
            self._allocate(self.capacity)

    # -- buffers ------------------------------------------------------------------------------
    def _agree(self, ok: bool) -> bool:
        """True only if every rank of the group says ok (ranks must take the same exchange path)."""
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda" if self.on_gpu else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(flag.item()))

    def _allocate(self, capacity: int):
        """capacity must be identical on every rank (symmetric buffers; callers derive it from gathered data)."""
        self.capacity = capacity
        if self.fused:
            err = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.recv = symm_mem.empty(capacity, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
            except Exception as e:  # symmetric memory unavailable on this rank
                err = repr(e)
            # a rank that could not allocate must not leave its peers waiting in the rendezvous / barriers
            if self._agree(err is None):
                try:
                    self.symm = symm_mem.rendezvous(self.recv, self.group.group_name)
                    self.peer_ptrs = [int(p) for p in self.symm.buffer_ptrs]
                except Exception as e:
                    err = repr(e)
                if not self._agree(err is None):
                    err = err or "a peer could not rendezvous"
            else:
                err = err or "a peer could not allocate symmetric memory"
            if err is not None:  # every rank falls back to the NCCL exchange together
                self.fused = False
                self.fused_error = err
                self.symm = None
                self.peer_ptrs = None
        if not self.fused:
            self.recv = self.ops.empty(capacity)
        self.out = self.ops.empty(capacity)
        self.recv_v = self.out_v = None      # value buffers are created by the first sort with values

    def _allocate_values(self):
        if self.recv_v is not None:
            return
        if self.fused:
            import torch.distributed._symmetric_memory as symm_mem
            self.recv_v = symm_mem.empty(self.capacity, dtype=torch.int32,
                                         device=torch.device("cuda", torch.cuda.current_device()))
            self.symm_v = symm_mem.rendezvous(self.recv_v, self.group.group_name)
            self.peer_ptrs_v = [int(p) for p in self.symm_v.buffer_ptrs]
        else:
            self.recv_v = self.ops.empty(self.capacity)
        self.out_v = self.ops.empty(self.capacity)

    # -- the sort -----------------------------------------------------------------------------
    def sort_pairs(self, keys, vals):
        """Stable key/value variant: returns (keys, values) slices of the globally sorted pairs."""
        return self.sort(keys, vals)

    def sort(self, keys, vals=None):
        """keys: this rank's shard (4-byte ints).  Returns this rank's slice of the globally sorted
        array (a view of an internal buffer, valid until the next call); with `vals`, a pair of
        slices (equal keys keep their global input order)."""
        ops, world, rank = self.ops, self.world, self.rank
        n_local = keys.numel()
        shift = 32 - TOP_BITS
        t = self.timer
        t.marks = []
        t.mark("start")

        # Partition digit = the highest 8-bit window in which the keys differ: when every key of every
        # rank shares the window's value (small-range keys: constant high bytes) the window moves down
        # one byte and the histogram is taken again -- the higher bytes are then equal for all keys, so
        # the lower window still defines a range partition.  Uniform keys never take a second round.
        while True:
            hist = ops.histogram(keys, shift, TOP_BITS)                 # device, uint32 counts as int32
            counts = hist.to(torch.int64) & 0xFFFFFFFF
            gathered = torch.empty(world * counts.numel(), dtype=torch.int64, device=counts.device)
            dist.all_gather_into_tensor(gathered, counts, group=self.group)
            counts_all = gathered.cpu().numpy().reshape(world, -1)      # sync point: split sizes live on the host
            if shift == 0 or np.count_nonzero(counts_all.sum(axis=0)) > 1:
                break
            shift -= TOP_BITS
        t.mark("histogram")
        self.partition_shift = shift
        if np.count_nonzero(counts_all.sum(axis=0)) <= 1:
            # all keys are equal: every shard already is a slice of the sorted array, in global input
            # order (which is what stability asks of equal keys) -- nothing to exchange
            plan = plan_exchange(counts_all, rank)
            plan["identity"] = True
            self.last_plan = plan
            if self.out is None or self.out.numel() < n_local:
                self.out = ops.empty(max(n_local, 1))
            self.out[:n_local].copy_(keys)
            t.mark("local_sort")
            if vals is None:
                return self.out[:n_local]
            if self.out_v is None or self.out_v.numel() < n_local:
                self.out_v = ops.empty(max(n_local, 1))
            self.out_v[:n_local].copy_(vals)
            return self.out[:n_local], self.out_v[:n_local]
        plan = plan_exchange(counts_all, rank)
        total = int(plan["totals"].sum())
        plan["imbalance"] = float(plan["totals"].max()) * world / max(total, 1)
        if world > 1 and self.balance_threshold > 0 and plan["imbalance"] > self.balance_threshold:
            return self._sort_by_value_splitters(keys, vals, counts_all.sum(axis=1))
        self.last_plan = plan
        self._ensure_capacity(plan, n_local, vals)
        t.mark("splitters")

        my_total = plan["my_total"]
        pbits = plan["narrow_bits"] if self.allow_narrow and plan["narrow_bits"] else TOP_BITS
        pshift = shift + TOP_BITS - pbits
        if self.fused:
            # partition + exchange in one kernel: every bin is stored straight into its owner's buffer
            if pbits == TOP_BITS:
                addr = np.array([self.peer_ptrs[int(o)] for o in plan["owner"]], dtype=np.int64) + 4 * plan["bin_recv_offset"]
            else:
                addr = np.array(self.peer_ptrs[:world], dtype=np.int64) + 4 * plan["src_base"]
            if vals is not None:
                delta = addr - np.array([self.peer_ptrs[int(o)] for o in plan["owner"]] if pbits == TOP_BITS
                                        else self.peer_ptrs[:world], dtype=np.int64)
                vaddr = np.array([self.peer_ptrs_v[int(o)] for o in plan["owner"]] if pbits == TOP_BITS
                                 else self.peer_ptrs_v[:world], dtype=np.int64) + delta
                addr = np.concatenate([addr, vaddr])
            bin_dst = torch.from_numpy(addr).to("cuda")
            self.symm.barrier(channel=0)        # peers are done reading their previous receive buffers
            ops.digit_pass(keys, pshift, pbits, bin_dst=bin_dst, vals=vals)
            self.symm.barrier(channel=1)        # every rank's stores have landed
            t.mark("partition+exchange")
        else:
            if self.part is None or self.part.numel() < n_local:
                self.part = ops.empty(n_local)
            if vals is None:
                part = ops.digit_pass(keys, pshift, pbits, out=self.part[:n_local])
            else:
                if self.part_v is None or self.part_v.numel() < n_local:
                    self.part_v = ops.empty(n_local)
                part, part_v = ops.digit_pass(keys, pshift, pbits, out=self.part[:n_local], vals=vals,
                                              out_vals=self.part_v[:n_local])
            t.mark("partition")
            self._all_to_all(self.recv[:my_total], part, plan["recv_counts"], plan["send_counts"])
            if vals is not None:
                self._all_to_all(self.recv_v[:my_total], part_v, plan["recv_counts"], plan["send_counts"])
            t.mark("exchange")
        return self._local_sort(my_total, vals is not None, key_bits=shard_key_bits(plan["owner"], rank, shift))

    def _ensure_capacity(self, plan, n_local, vals):
        need = int(plan["totals"].max())
        if self.recv is None or need > self.capacity:
            if self.fused and self.recv is not None:
                raise RuntimeError(f"receive capacity {self.capacity} < {need}: construct ShardedSorter with a larger per_rank_capacity")
            # sized from values every rank holds (the gathered counts), never from the local shard alone
            largest_shard = int(plan["shard_sizes"].max()) if "shard_sizes" in plan else n_local
            self._allocate(max(need, int(largest_shard * 1.05) + 1024))
        if vals is not None:
            self._allocate_values()

    def _local_sort(self, my_total, pairs, key_bits=32):
        """key_bits < 32: the shard's keys agree in their bits >= key_bits (the splitters fixed them), so the local
        sort skips those digits (b200sort_keys_low_bits)."""
        ops, t = self.ops, self.timer
        out = self.out[:my_total]
        # worth it only while the number of digit passes stays the same (the narrower top digit runs a smaller
        # kernel): with fewer passes the histogram kernel leaves its compile-time form and costs more than a pass
        if -(-key_bits // self.nbits) != -(-32 // self.nbits):
            key_bits = 32
        kw = {"key_bits": key_bits} if key_bits < 32 else {}
        self.last_key_bits = key_bits
        if not pairs:
            if my_total:
                ops.sort(self.recv[:my_total], self.nbits, out, **kw)
            t.mark("local_sort")
            return out
        out_v = self.out_v[:my_total]
        if my_total:
            ops.sort(self.recv[:my_total], self.nbits, out, vals=self.recv_v[:my_total], out_vals=out_v, **kw)
        t.mark("local_sort")
        return out, out_v

    def _sort_by_value_splitters(self, keys, vals, shard_sizes):
        """Skewed keys: the edges of the 256 bins of the partition byte cannot balance the shards
        (a heavy bin is never split).  Splitters become key VALUES taken from a sample of every
        shard; each key's destination (b200sort_route) is then the key of a stable digit pass that
        carries the real keys -- and, in a second pass, the values -- to their owners.  Costs one
        extra read+write of the shard.  A run of equal keys that straddles a cut is split at a
        position of the global input order (value_splitters), which keeps ties in that order."""
        ops, world, rank, t = self.ops, self.world, self.rank, self.timer
        n_local = keys.numel()
        m = SAMPLE_PER_RANK
        idx = sample_indices(n_local, m)
        mine = ops.sample(keys, idx) if n_local else ops.empty(m).zero_()
        gathered = ops.empty(world * m)
        dist.all_gather_into_tensor(gathered, mine.contiguous(), group=self.group)
        valid = [r for r in range(world) if int(shard_sizes[r]) > 0]
        pool = gathered if len(valid) == world else torch.cat([gathered[r * m:(r + 1) * m] for r in valid])
        pool_sorted = ops.sort(pool, 8, ops.empty(pool.numel()))
        t.mark("splitters:sample")
        by_rank = gathered.cpu().numpy().view(np.uint32).reshape(world, m)
        empty = np.zeros(0, dtype=np.uint32)
        values, split_rank, split_pos = value_splitters(
            pool_sorted.cpu().numpy().view(np.uint32), [by_rank[r] if r in valid else empty for r in range(world)],
            [sample_indices(int(shard_sizes[r]), m) for r in range(world)], world)
        cut_values, cut_ties = thresholds_for_rank(values, split_rank, split_pos, rank)     # this rank's own cuts
        t.mark("splitters:cuts")

        bits = max(1, (world - 1).bit_length())
        route, dest_counts = ops.route(keys, cut_values, cut_ties)      # destinations + how many keys go to each
        t.mark("splitters:route")
        counts = dest_counts.to(torch.int64) & 0xFFFFFFFF
        allc = torch.empty(world * counts.numel(), dtype=torch.int64, device=counts.device)
        dist.all_gather_into_tensor(allc, counts, group=self.group)
        counts_all = np.zeros((world, 1 << bits), dtype=np.int64)      # bins >= world stay empty
        counts_all[:, :world] = allc.cpu().numpy().reshape(world, world)
        owner = np.minimum(np.arange(1 << bits), world - 1)
        plan = plan_exchange(counts_all, rank, owner=owner)
        plan["narrow_bits"] = 0
        plan["value_thresholds"] = (cut_values, cut_ties)
        total = int(plan["totals"].sum())
        plan["imbalance"] = float(plan["totals"].max()) * world / max(total, 1)
        self.last_plan = plan
        self.partition_bits = bits
        self._ensure_capacity(plan, n_local, vals)
        if self.part is None or self.part.numel() < n_local:
            self.part = ops.empty(max(n_local, 1))
        if self.route_dump is None or self.route_dump.numel() < n_local:
            self.route_dump = ops.empty(max(n_local, 1))
        t.mark("splitters")

        my_total = plan["my_total"]
        if self.fused:
            # the route array itself stays local (dump), the carried array goes to the peers
            local_off = np.concatenate([[0], np.cumsum(counts_all[rank])[:-1]]).astype(np.int64)
            dump_addr = self.route_dump.data_ptr() + 4 * local_off
            peer = np.array([self.peer_ptrs[int(o)] for o in owner], dtype=np.int64) + 4 * plan["bin_recv_offset"]
            self.symm.barrier(channel=0)
            ops.digit_pass(route, 0, bits, bin_dst=torch.from_numpy(np.concatenate([dump_addr, peer])).to("cuda"), vals=keys)
            if vals is not None:
                peer_v = np.array([self.peer_ptrs_v[int(o)] for o in owner], dtype=np.int64) + 4 * plan["bin_recv_offset"]
                ops.digit_pass(route, 0, bits, bin_dst=torch.from_numpy(np.concatenate([dump_addr, peer_v])).to("cuda"), vals=vals)
            self.symm.barrier(channel=1)
            t.mark("partition+exchange")
        else:
            dump = self.route_dump[:n_local]
            _, part = ops.digit_pass(route, 0, bits, out=dump, vals=keys, out_vals=self.part[:n_local])
            t.mark("partition")
            self._all_to_all(self.recv[:my_total], part, plan["recv_counts"], plan["send_counts"])
            if vals is not None:
                if self.part_v is None or self.part_v.numel() < n_local:
                    self.part_v = ops.empty(max(n_local, 1))
                _, part_v = ops.digit_pass(route, 0, bits, out=dump, vals=vals, out_vals=self.part_v[:n_local])
                self._all_to_all(self.recv_v[:my_total], part_v, plan["recv_counts"], plan["send_counts"])
            t.mark("exchange")
        return self._local_sort(my_total, vals is not None)

    def _all_to_all(self, recv, send, recv_counts, send_counts):
        rc = [int(c) for c in recv_counts]
        sc = [int(c) for c in send_counts]
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=self.group)
            return
        # gloo (CPU tests): point-to-point emulation of the same exchange
        ro = np.concatenate([[0], np.cumsum(rc)]).astype(np.int64)
        so = np.concatenate([[0], np.cumsum(sc)]).astype(np.int64)
        ops_list = []
        for peer in range(self.world):
            if peer == self.rank:
                recv[ro[peer]:ro[peer + 1]].copy_(send[so[peer]:so[peer + 1]])
                continue
            if sc[peer]:
                ops_list.append(dist.P2POp(dist.isend, send[so[peer]:so[peer + 1]].contiguous(),
                                           dist.get_global_rank(self.group, peer), group=self.group))
            if rc[peer]:
                ops_list.append(dist.P2POp(dist.irecv, recv[ro[peer]:ro[peer + 1]],
                                           dist.get_global_rank(self.group, peer), group=self.group))
        if ops_list:
            for req in dist.batch_isend_irecv(ops_list):
                req.wait()

    # -- reporting ----------------------------------------------------------------------------
    def phase_report(self) -> dict:
        """Per-phase milliseconds of the LAST sort, max over ranks; plus exchange bytes / NVLink figure."""
        local = self.timer.report()
        if not local:
            return {}
        names = sorted(local)
        vals = torch.tensor([local[k] for k in names], dtype=torch.float64, device="cuda")
        dist.all_reduce(vals, op=dist.ReduceOp.MAX, group=self.group)
        rep = {k: float(v) for k, v in zip(names, vals.tolist())}
        plan = self.last_plan
        if plan is not None:
            sent = int(plan["send_counts"].sum() - plan["send_counts"][self.rank]) * 4
            rep["egress_bytes_this_rank"] = sent
            key = "partition+exchange" if self.fused else "exchange"
            if rep.get(key):
                rep["egress_gbs_this_rank"] = sent / (rep[key] * 1e-3) / 1e9
            rep["exchange"] = "fused peer stores" if self.fused else "nccl all_to_all_single"
            rep["partition_bits"] = int(plan["narrow_bits"]) if self.allow_narrow and plan["narrow_bits"] else TOP_BITS
            if "value_thresholds" in plan:
                rep["partition_bits"] = int(self.partition_bits)
                rep["splitter_kind"] = "values from a sample"
            rep.setdefault("splitter_kind", "bin edges of the partition byte")
            rep["imbalance"] = float(plan.get("imbalance", 0.0))
            rep["shard_sizes"] = [int(x) for x in plan["totals"]]
            rep["partition_shift"] = int(getattr(self, "partition_shift", 32 - TOP_BITS))
        return rep


# ---------------------------------------------------------------------------------------------
def verify_sharded(result, keys, group=None, verify_fn=None) -> bool:
    """Size-independent check of a sharded sort (used at 2^32 keys where no host can hold the
    oracle's answer): every shard non-decreasing, shard boundaries ordered, total count kept and
    the order-independent multiset fingerprint (sum key, sum sm64(key), xor sm64(key)) unchanged."""
    group = group if group is not None else dist.group.WORLD
    world = dist.get_world_size(group)
    if verify_fn is None:
        from . import api
        verify_fn = api.verify
    dev = result.device
    bad, s1, h1, x1 = verify_fn(result) if result.numel() else (0, 0, 0, 0)
    _, s0, h0, x0 = verify_fn(keys) if keys.numel() else (0, 0, 0, 0)
    lo = int(result[0].item()) & 0xFFFFFFFF if result.numel() else -1
    hi = int(result[-1].item()) & 0xFFFFFFFF if result.numel() else -1

    def i64(v):  # two's-complement wrap into int64
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v

    mine = torch.tensor([bad, result.numel(), keys.numel(), i64(s1), i64(h1), i64(x1), i64(s0), i64(h0), i64(x0), lo, hi],
                        dtype=torch.int64, device=dev)
    allv = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    rows = allv.cpu().numpy().reshape(world, -1)
    ok = int(rows[:, 0].sum()) == 0 and int(rows[:, 1].sum()) == int(rows[:, 2].sum())
    m = (1 << 64) - 1
    ok &= (sum(int(v) for v in rows[:, 3]) & m) == (sum(int(v) for v in rows[:, 6]) & m)
    ok &= (sum(int(v) for v in rows[:, 4]) & m) == (sum(int(v) for v in rows[:, 7]) & m)
    x_out = x_in = 0
    for r in range(world):
        x_out ^= int(rows[r, 5]) & m
        x_in ^= int(rows[r, 8]) & m
    ok &= x_out == x_in
    prev_hi = -1
    for r in range(world):
        if rows[r, 1] == 0:
            continue
        ok &= int(rows[r, 9]) >= prev_hi
        prev_hi = int(rows[r, 10])
    return bool(ok)
