#!/usr/bin/env python
"""NVLink store bandwidth from SM-issued stores into a peer's symmetric-memory buffer (torchrun, >= 2 ranks)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
import cuda.radixsort_b200 as rs
from cuda.radixsort_b200 import _lib

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = rs.load()
    n = 1 << 28
    buf = symm_mem.empty(n, dtype=torch.int32, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
    src = torch.arange(n, dtype=torch.int32, device="cuda")
    peer = (rank + 1) % world
    out = {}
    for target, name in ((hdl.buffer_ptrs[peer], "peer"), (hdl.buffer_ptrs[rank], "local")):
        for vec in (1, 4):
            for ctas in (2, 4, 8):
                hdl.barrier(channel=0)
                for _ in range(2):
                    _lib.check(lib.b200sort_store_probe(target, src.data_ptr(), n, vec, ctas, torch.cuda.current_stream().cuda_stream))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    _lib.check(lib.b200sort_store_probe(target, src.data_ptr(), n, vec, ctas, torch.cuda.current_stream().cuda_stream))
                b.record(); torch.cuda.synchronize()
                out[f"{name}_vec{vec}_ctas{ctas}_GBps_written"] = round(4 * n / (a.elapsed_time(b) / 5) / 1e6, 1)
                hdl.barrier(channel=1)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
