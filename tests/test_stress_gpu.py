"""Stress / litmus tests of the two protocols whose correctness rests on more than data flow (stand-ins for
compute-sanitizer racecheck, which is closed on this pool -- SURVEY section 5 "race detection", section 7.2
"look-back forward progress"):

  * the decoupled look-back (descriptor flags with relaxed gpu-scope accesses, parity-rotated status codes,
    tile id = blockIdx, forward progress from in-order CTA dispatch): ~10^4 launches of the digit-pass kernels
    at tiny tiles -- many tiles in flight, every look-back many descriptors deep -- on four streams at once,
    each launch checked (sortedness + multiset fingerprint + equality across repetitions);
  * the ranking atomics: the on-device self test under interference, on every visible device.

Replaces nothing in the reference (it has one stream and a device synchronisation after every kernel,
Parallel7.cu:224-233); the property tested is the one its element-wise check states (Parallel7.cu:679-687).
"""
import numpy as np
import pytest

from conftest import to_dev, to_host

pytestmark = pytest.mark.gpu


def _litmus_variants():
    """Column sweep (36) and the default (95, two chains + vector look-back); with the tuning build also tiny tiles
    (28, ballot rank: many tiles in flight) and atomic rank (10)."""
    import os
    return [28, 10, 36, 95] if os.environ.get("B200_TUNING") else [36, 95]


@pytest.mark.parametrize("variant", _litmus_variants())
def test_look_back_litmus_many_small_launches_on_concurrent_streams(rs, oracle, variant):
    import torch
    rs.set_param("variant", variant)
    try:
        tile = rs.tile_keys(False)
        streams = [torch.cuda.Stream() for _ in range(4)]
        launches = 0
        # sizes: a few hundred tiles per launch, ragged last tile, plus one single-tile and one two-tile case
        sizes = [tile * 200 + 77, tile * 333 + 1, tile * 64 - 5, tile + 1, tile - 1]
        rounds = 60 if variant in (28, 36) else 24
        for n in sizes:
            k = oracle.generate("zipf" if n % 2 else "uniform", n)
            want = oracle.sort_keys(k, 8)
            d = to_dev(k)
            _, s0, h0, x0 = rs.verify(d)
            outs = [torch.empty_like(d) for _ in streams]
            wss = [rs.Workspace("cuda") for _ in streams]
            for ws in wss:
                ws.get(rs.temp_bytes(n, 8, False))
            torch.cuda.synchronize()
            for r in range(rounds):
                for st, o, ws in zip(streams, outs, wss):
                    with torch.cuda.stream(st):
                        rs.sort_keys(d, 8, out=o, workspace=ws, stream=st)      # 5 launches: hist + 4 digit passes
                        launches += 5
                if r % 6 == 5 or r == rounds - 1:
                    torch.cuda.synchronize()
                    for o in outs:
                        bad, s1, h1, x1 = rs.verify(o)
                        assert bad == 0 and (s1, h1, x1) == (s0, h0, x0), (variant, n, r)
            torch.cuda.synchronize()
            for o in outs:
                assert np.array_equal(to_host(o), want), (variant, n)
        assert launches >= (5000 if variant in (28, 36) else 2000)
    finally:
        rs.set_param("variant", -1)


def test_look_back_litmus_one_bit_digits_32_passes(rs, oracle):
    """32 digit passes per sort (nBits=1): the descriptor array is cleared once and reused by 32 launches
    with alternating status-code parity."""
    import torch
    n = (1 << 20) + 3
    k = oracle.generate("uniform", n)
    want = oracle.sort_keys(k, 1)
    d = to_dev(k)
    for _ in range(20):
        assert np.array_equal(to_host(rs.sort_keys(d, 1)), want)
    torch.cuda.synchronize()


def test_atomic_order_selftest_runs_on_every_visible_device(rs):
    import torch
    verdicts = []
    for dev in range(torch.cuda.device_count()):
        with torch.cuda.device(dev):
            verdicts.append(rs.get_param("atomic_rank_ok"))
    assert all(v in (0, 1) for v in verdicts)
    with torch.cuda.device(0):
        # the default kernel does not depend on the verdict (column sweep with two chains, mode 4) ...
        assert rs.get_param("rank_mode") == 4
        # ... and an atomic-rank request (mode 1) is honoured only on a device that passed
        rs.set_param("variant", 10)
        try:
            if rs.get_param("tuning_build"):
                assert rs.get_param("rank_mode") == (1 if verdicts[0] == 1 else 3)
            else:
                assert rs.get_param("rank_mode") == 4   # not in the product library: the default kernel
        finally:
            rs.set_param("variant", -1)
