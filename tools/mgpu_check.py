#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
Bit-exact against the oracle at a reduced total size, for the NCCL exchange and the fused
peer-store exchange, then the property check at a larger size."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cuda.radixsort_b200 as rs  # noqa: E402
from cuda.radixsort_b200 import mgpu  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import oracle as O
    report = {"world": world}
    for fused in (False, True):
        for kind, log2n in (("uniform", 24), ("zipf", 22), ("unique16", 20), ("all_equal", 18), ("iota", 20),
                            ("heavybin", 22), ("heavyvalue", 22)):
            total = (1 << log2n) + 12345
            per = total // world
            first = rank * per
            count = per if rank < world - 1 else total - first
            cdf = O.zipf_cdf() if kind == "zipf" else None
            if kind == "heavybin":   # 90 % of the keys share one top byte and differ below it
                keys = rs.generate("uniform", count, first=first, total=total)
                pick = rs.generate("uniform", count, first=first + total, total=2 * total)
                heavy = torch.remainder(pick.to(torch.int64) & 0xFFFFFFFF, 10) != 0
                keys = torch.where(heavy, (keys & 0x00FFFFFF) | 0x5A000000, keys)
            elif kind == "heavyvalue":   # 70 % of the keys are ONE value: its run is cut between ranks at a position
                keys = rs.generate("uniform", count, first=first, total=total)
                pick = rs.generate("uniform", count, first=first + total, total=2 * total)
                heavy = torch.remainder(pick.to(torch.int64) & 0xFFFFFFFF, 10) < 7
                keys = torch.where(heavy, torch.full_like(keys, 0x5A5A5A5A), keys)
            else:
                keys = rs.generate(kind, count, first=first, total=total, zipf_cdf=cdf)
            sorter = mgpu.ShardedSorter(dist.group.WORLD, per_rank_capacity=total + 1024, nbits=8, fused=fused)
            res = sorter.sort(keys)
            res2 = sorter.sort(keys)                      # buffers are reused correctly
            assert torch.equal(res, res2)
            ok = mgpu.verify_sharded(res, keys)
            # key/value variant: value = global index; equal keys must keep global input order
            vals = torch.arange(first, first + count, dtype=torch.int32, device="cuda")
            res = res.clone()
            pk, pv = sorter.sort_pairs(keys, vals)
            assert torch.equal(pk, res), "pair sort disagrees with key sort"
            same = pk[1:] == pk[:-1]
            assert bool(torch.all(pv[1:][same] > pv[:-1][same])), "equal keys out of input order"
            # every (key, value) pair is an input pair: gather the whole input and look the values up
            full = torch.zeros(total, dtype=torch.int32, device="cuda")
            full[first:first + count] = keys
            dist.all_reduce(full)                           # disjoint slices: sum == concatenation
            ok = ok and bool(torch.equal(full[pv.long()], pk))
            sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
            sizes[rank] = res.numel()
            dist.all_reduce(sizes)
            pad = torch.zeros(total, dtype=torch.int32, device="cuda")
            off = int(sizes[:rank].sum().item())
            pad[off:off + res.numel()] = res
            dist.all_reduce(pad)                          # disjoint slices: sum == concatenation
            if rank == 0:
                whole = full.cpu().numpy().view(np.uint32)
                exact = bool(np.array_equal(pad.cpu().numpy().view(np.uint32), O.sort_keys(whole, 8)))
                report[f"{'fused' if sorter.fused else 'nccl'}:{kind}:2^{log2n}"] = {
                    "bit_exact": exact, "verify": ok, "shards": [int(x) for x in sizes.tolist()],
                    "splitters": "values" if "value_thresholds" in sorter.last_plan else "bin edges",
                    "imbalance": round(float(sorter.last_plan.get("imbalance", 0.0)), 4),
                    "requested_fused": fused, "fused_error": getattr(sorter, "fused_error", None)}
                assert exact and ok, (kind, fused)
            del sorter
    if rank == 0:
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
