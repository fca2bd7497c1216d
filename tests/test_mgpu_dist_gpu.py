"""The sharded multi-GPU sort -- the exact code bench.py --gpus N times (mgpu.ShardedSorter, one
process per GPU under torch.distributed.run, NCCL) -- against the oracle, bit-exact, on the real
GPUs of the box: both exchange modes (fused peer stores / NCCL all_to_all) x uniform, Zipf, 16
values, all-equal, iota, 90 %-in-one-bin and 70 %-one-value inputs, keys and stable pairs
(tools/mgpu_check.py; SURVEY section 8e "parity at scale": concatenated shards == sortByHost of the
whole input, Baseline1.cu:15-64).  Needs >= 2 visible GPUs; ranks must never share a GPU (kernels
of different ranks wait on one another).
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_sorter_bit_exact_on_all_visible_gpus(rs):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (one rank per GPU)")
    if world not in (2, 4, 8):
        world = 4 if world > 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    report = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert report["world"] == world
    cases = {k: v for k, v in report.items() if isinstance(v, dict)}
    assert len(cases) >= 14
    assert all(v["bit_exact"] and v["verify"] for v in cases.values()), cases
