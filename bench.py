#!/usr/bin/env python
"""bench.py -- headline benchmark: uint32 keys/s sorted (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload uniform|zipf|unique16|all_equal|sorted|reversed|pairs] [--log2n L]

N = 1: a step is one LSD sort of 2^28 uniform uint32 keys with 8-bit digits (BASELINE config
1) through libb200sort.so.  `value` = keys/s with the keys resident in HBM (CUDA events on
the launching stream); `e2e` = the same sort through the reference-facing host entry point
sort(in, n, out, SORT_BY_DEVICE, nBits, blockSize) with pinned HOST buffers, H2D + D2H inside
the timed region.  `roofline` is for the dominant kernel (the digit-pass kernel: 8 bytes per
key per launch), timed live with CUDA events recorded by the library on the sort's stream.
`cpu_baseline` times the reference's own sortByHost (oracle/_ref when present, else the C
port) on a bounded sample on the box's host cores.

N > 1 (torchrun, one rank per GPU): a step is one sharded sort of 2^32 keys -- MSD partition,
NVLink exchange, local LSD sort (cuda/radixsort_b200/mgpu.py).

--impl reference: the reference's CPU sort only (rank 0), on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "uint32 keys/sec sorted"
UNIT = "keys/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, "of fallback"


# --------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic(kernel: str):
    """Per-launch DRAM bytes of `kernel` from the committed ncu summary, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def committed_traffic_source():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("_source")
    except Exception:
        return None


def _run_json(cmd, timeout):
    """Runs a helper process and returns the last JSON line it printed (or an 'unavailable' note)."""
    try:
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"unavailable": f"rc={r.returncode}: {(r.stderr or r.stdout)[-200:]}"}
    except Exception as e:  # missing binary, timeout
        return {"unavailable": repr(e)[:200]}


def same_box_gpu_baselines(log2n: int, nbits: int):
    """The two GPU bars BASELINE.md promises beside our number, timed in this run on this GPU, each in
    its own process: (1) the reference's best version as shipped -- Parallel7 sortByDevice, unmodified
    (oracle/_ref; Parallel7.cu:530-639), host arrays in and out; (2) device-resident
    cub::DeviceRadixSort, what the reference's sortByThrust (Baseline1.cu:66-70) resolves to
    (tools/cub_bench.cu, a bench-only binary: libb200sort.so links neither CUB nor Thrust)."""
    ref = _run_json([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_time.py"), "--log2n", str(log2n),
                     "--nbits", str(nbits), "--block", "512"], timeout=300)
    cub_bin = os.path.join(ROOT, "tools", "_bin", "cub_bench")
    cub = _run_json([cub_bin, str(log2n), "10"], timeout=300) if os.path.exists(cub_bin) else \
        {"unavailable": "tools/_bin/cub_bench not built (__graft_entry__.build())"}
    return ref, cub


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
def cpu_reference_sort(sample_log2: int, nbits: int, reps: int, kind_pref: str = "auto"):
    """Times the reference's sortByHost (single-threaded by construction, Baseline1.cu:30-55)
    on 2^sample_log2 uniform keys.  Returns (keys_per_s, kind, seconds_per_sort)."""
    import oracle as O
    n = 1 << sample_log2
    keys = O.generate("uniform", n)
    use_ref = kind_pref != "port" and O.ref_available("Baseline1")
    fn = (lambda: O.ref_sort_by_host(keys, nbits)) if use_ref else (lambda: O.sort_keys(keys, nbits))
    times = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - t0)
    assert O.is_sorted(out)
    best = min(times)
    return n / best, ("reference" if use_ref else "port"), best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # One core sorts ~45 Mkeys/s: the full 2^28-key workload takes ~6 s per step.  Sort the REAL n when
    # the whole run stays within a few minutes, else a bounded sample (and say which: `sample_n`).
    sample_log2 = args.cpu_sample_log2
    if args.gpus == 1 and (args.steps + args.warmup) * (1 << args.log2n) * 22e-9 <= 240.0:
        sample_log2 = args.log2n
    times = []
    import oracle as O
    n = 1 << sample_log2
    keys = O.generate("uniform", n)
    use_ref = O.ref_available("Baseline1")
    fn = (lambda: O.ref_sort_by_host(keys, args.nbits)) if use_ref else (lambda: O.sort_keys(keys, args.nbits))
    for _ in range(args.warmup):
        fn()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = fn()
        times.append(time.perf_counter() - t0)
    assert O.is_sorted(out)
    total = sum(times)
    value = n * args.steps / total
    workload_log2 = args.log2n if args.gpus == 1 else args.log2n_multi
    sample = (f"2^{sample_log2} uniform uint32 keys per step ("
              + ("the full workload" if sample_log2 == workload_log2 else
                 f"a bounded sample of the 2^{workload_log2}-key workload: keys/s of a counting sort does not "
                 f"grow with n, and sortByHost takes an int n")
              + f"), sortByHost nBits={args.nbits}, single thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1,
                         "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "sample_n": n, "workload_n": 1 << workload_log2,
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, n_gpus: int) -> dict:
    if n_gpus == 1:
        what = "key+value pairs" if args.workload == "pairs" else "keys"
        dist = "uniform" if args.workload == "pairs" else args.workload
        return {"workload": f"2^{args.log2n} {dist} uint32 {what}, {args.nbits}-bit digits "
                            f"({-(-32 // args.nbits)} passes), 1 B200",
                "n": 1 << args.log2n, "nbits": args.nbits, "distribution": dist,
                "l2": "inputs larger than L2 (1 GiB of keys vs 126 MB), no flush needed"}
    total_log2 = args.log2n_multi
    what = "key+value pairs" if args.workload == "pairs" else "keys"
    dist = "uniform" if args.workload == "pairs" else args.workload
    return {"workload": f"2^{total_log2} {dist} uint32 {what} sharded over {n_gpus} B200: top-digit "
                        f"histogram + all-gather splitters, MSD partition, NVLink all-to-all, local LSD sort",
            "n": 1 << total_log2, "nbits": args.nbits, "distribution": dist,
            "l2": "inputs larger than L2"}


# --------------------------------------------------------------------------------------------
def run_single(args):
    import torch

    import cuda.radixsort_b200 as rs

    torch.cuda.set_device(0)
    rs.load()
    n = 1 << args.log2n
    nbits = args.nbits
    gpu_reference = cub = None
    if not args.no_gpu_baselines:
        gpu_reference, cub = same_box_gpu_baselines(args.log2n, nbits)
    pairs = args.workload == "pairs"
    dist = "uniform" if pairs else args.workload
    cdf = None
    if dist == "zipf":
        import oracle as O          # generator table only (test infrastructure; not the sort)
        cdf = O.zipf_cdf()
    keys = rs.generate(dist, n, zipf_cdf=cdf)
    out = torch.empty_like(keys)
    vals = vout = None
    if pairs:
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        vout = torch.empty_like(vals)
    ws = rs.Workspace("cuda")
    ws.get(rs.temp_bytes(n, nbits, pairs))

    def step():
        if pairs:
            rs.sort_pairs(keys, vals, nbits, out_keys=out, out_vals=vout, workspace=ws)
        else:
            rs.sort_keys(keys, nbits, out=out, workspace=ws)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    rs.profile_enable(True)
    rs.profile_read()
    launches0 = rs.launch_count()
    sampler = ClockSampler(0).start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = rs.launch_count() - launches0
    total_ms = start.elapsed_time(stop)
    prof = rs.profile_read()
    rs.profile_enable(False)
    ms_per_step = total_ms / args.steps
    value = n / (ms_per_step * 1e-3)

    # correctness of what was just timed: sortedness + multiset fingerprint on the device
    bad, s1, h1, x1 = rs.verify(out)
    _, s0, h0, x0 = rs.verify(keys)
    assert bad == 0 and (s1, h1, x1) == (s0, h0, x0), "timed output is not a sorted permutation of the input"

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak, peak_src = measured_peak()
    pass_ms = [ms for tag, ms in prof if tag >= 1]
    hist_ms = [ms for tag, ms in prof if tag == 0]
    bytes_per_key = 16 if pairs else 8
    pass_avg = sum(pass_ms) / max(1, len(pass_ms))
    achieved = bytes_per_key * n / (pass_avg * 1e-3) / 1e9 if pass_ms else None
    eff_variant = rs.get_param("effective_variant") if not pairs else None
    # the digit-pass kernel in effect: the column sweep (csrc/colsweep.cuh, rank modes 3 and 4) or onesweep.cuh
    kernel_name = "colsweep_pass_kernel" if rs.get_param("rank_mode") >= 3 else "onesweep_pass_kernel"
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None,
        "traffic": committed_traffic(kernel_name),
        "traffic_source": "NOT measured in this run: " + str(committed_traffic_source()),
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": bytes_per_key * n,
        "avg_launch_ms": pass_avg, "launches_timed": len(pass_ms),
        "hist_kernel_avg_ms": (sum(hist_ms) / len(hist_ms)) if hist_ms else None,
        "kernel_share_of_step": (sum(pass_ms) / total_ms) if pass_ms else None,
        "whole_sort": {"algorithmic_bytes": rs.algorithmic_bytes(n, nbits, pairs),
                       "achieved": rs.algorithmic_bytes(n, nbits, pairs) / (ms_per_step * 1e-3) / 1e9,
                       "frac": rs.algorithmic_bytes(n, nbits, pairs) / (ms_per_step * 1e-3) / 1e9 / peak},
    }

    # ---- e2e: host buffers through the reference-facing entry point -------------------------
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n, dtype=torch.int32).pin_memory()
        h_out = torch.empty(n, dtype=torch.int32).pin_memory()
        h_in.copy_(keys)
        torch.cuda.synchronize()
        a_in, a_out = h_in.numpy().view(np.uint32), h_out.numpy().view(np.uint32)
        if pairs:
            hv_in = torch.arange(n, dtype=torch.int32).pin_memory()
            hv_out = torch.empty(n, dtype=torch.int32).pin_memory()
            b_in, b_out = hv_in.numpy().view(np.uint32), hv_out.numpy().view(np.uint32)
            call = lambda: rs.sort_pairs_by_device(a_in, b_in, n, a_out, b_out, nbits, 512)  # noqa: E731
        else:
            call = lambda: rs.sort(a_in, n, a_out, rs.SORT_BY_DEVICE, nbits, 512)            # noqa: E731
        for _ in range(max(1, min(args.warmup, 2))):
            call()
        e2e_steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        dt = (time.perf_counter() - t0) / e2e_steps
        # the host-path output is checked like the device-resident one: inversions + multiset fingerprint
        chk = torch.empty_like(keys)
        chk.copy_(h_out)
        bad_e, s_e, h_e, x_e = rs.verify(chk)
        assert bad_e == 0 and (s_e, h_e, x_e) == (s0, h0, x0), "e2e output is not a sorted permutation of the input"
        if pairs:
            chk.copy_(hv_out)
            assert torch.equal(keys[chk.to(torch.int64) & 0xFFFFFFFF], out), "e2e values do not follow their keys"
        del chk
        per = 8 * n if pairs else 4 * n
        e2e = {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": per, "d2h_bytes_per_step": per,
               "ms_per_step": dt * 1e3, "steps": e2e_steps, "verified": True,
               "how": "blocking host-pointer call sort(in,n,out,SORT_BY_DEVICE,nBits,512) on pinned "
                      "host buffers; wall clock around the calls (H2D + sort + D2H inside)"}
        if not pairs:
            # the reference's own caller passes malloc'ed (pageable) arrays (Parallel7.cu:712-715): same call
            p_in = np.array(a_in, copy=True)
            p_out = np.empty_like(p_in)
            rs.sort(p_in, n, p_out, rs.SORT_BY_DEVICE, nbits, 512)
            t0 = time.perf_counter()
            for _ in range(2):
                rs.sort(p_in, n, p_out, rs.SORT_BY_DEVICE, nbits, 512)
            dtp = (time.perf_counter() - t0) / 2
            assert np.array_equal(p_out[:: 1 << 12], a_out[:: 1 << 12])
            e2e["pageable"] = {"value": n / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3, "steps": 2,
                               "how": "same call on pageable (numpy-owned) host arrays, staged through the "
                                      "library's pinned ring by its host copy threads"}
            del p_in, p_out
        del h_in, h_out

    # ---- CPU baseline (reported, not the target) ---------------------------------------------
    cpu = None
    if not args.no_cpu:
        v, kind, secs = cpu_reference_sort(args.cpu_sample_log2, nbits, args.cpu_reps)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"2^{args.cpu_sample_log2} uniform uint32 keys (bounded sample of the 2^{args.log2n}-key "
                         f"workload), best of {args.cpu_reps}, sortByHost nBits={nbits}, {secs:.2f} s per sort; "
                         f"host has {os.cpu_count()} logical cores, the reference CPU sort is single-threaded"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args, 1),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "variant": rs.get_param("variant"), "tile_keys": rs.tile_keys(pairs),
        "effective_variant": eff_variant, "atomic_rank_ok": rs.get_param("atomic_rank_ok"),
        "gpu_reference": gpu_reference, "cub": cub,
        "parity_note": ("pairs: parity unpinned by the reference (it has no key/value path); checked against the "
                        "payload-carrying restatement of Baseline1's counting sort") if pairs else None,
    }
    if gpu_reference and "ms" in gpu_reference and e2e:
        line["vs_gpu_reference_e2e"] = gpu_reference["ms"] / e2e["ms_per_step"]
    if cub and "sortkeys_ms" in cub:
        line["vs_cub_device_resident"] = (cub["sortpairs_ms"] if pairs else cub["sortkeys_ms"]) / ms_per_step
    print(json.dumps(line), flush=True)
    return 0


def run_primitive(args):
    """--workload scan|hist: the two reusable primitives the reference treats as stages of their own
    (Docs/Snippets/PrefixSum-WorkEfficient.cu:229 scanByDevice, Docs/Snippets/Histogram.cu:17), device-resident,
    CUDA events, input larger than L2 when log2n >= 26 (else L2 is flushed between steps by a 256 MiB write)."""
    import torch

    import cuda.radixsort_b200 as rs

    torch.cuda.set_device(0)
    rs.load()
    n = 1 << args.log2n
    what = args.workload
    ref = None
    if not args.no_gpu_baselines:
        ref = _run_json([sys.executable, os.path.join(ROOT, "tools", "ref_primitive_time.py"), "--what", what,
                         "--log2n", str(min(args.log2n, 24))], timeout=300)
    keys = rs.generate("uniform", n)
    if what == "scan":
        x = (keys & 3).contiguous()                          # rand() & 0b11 like the snippet's main()
        out = torch.empty_like(x)
        ws = rs.Workspace("cuda")
        step = lambda: rs.exclusive_scan(x, out=out, workspace=ws)      # noqa: E731
        bytes_per_elem, kernel = 8, "exclusive_scan_kernel"
    else:
        x = keys
        ws = rs.Workspace("cuda")
        holder = {}

        def step():
            holder["h"] = rs.histogram(x, 0, 8, workspace=ws)
        bytes_per_elem, kernel = 4, "hist_kernel"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if n * 4 < (512 << 20) else None
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = rs.launch_count()
    sampler = ClockSampler(0).start()
    times = []
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    clocks = sampler.stop()
    launches = rs.launch_count() - launches0
    ms = sum(times) / len(times)
    # correctness of what was timed
    if what == "scan":
        want = torch.cumsum(x.to(torch.int64), 0) - x.to(torch.int64)
        assert torch.equal(out.to(torch.int64) & 0xFFFFFFFF, want & 0xFFFFFFFF), "scan differs from torch.cumsum"
    else:
        want = torch.bincount((x & 255).to(torch.int64), minlength=256)
        assert torch.equal(holder["h"].to(torch.int64), want), "histogram differs from torch.bincount"
    peak, peak_src = measured_peak()
    achieved = bytes_per_elem * n / (ms * 1e-3) / 1e9
    line = {
        "metric": f"uint32 elements/sec ({'exclusive scan' if what == 'scan' else '256-bin histogram'})",
        "value": n / (ms * 1e-3), "unit": "elements/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"{what} of 2^{args.log2n} uint32, device-resident, 1 B200", "n": n,
                   "l2": "input larger than L2" if flush is None else "L2 flushed between steps (256 MiB write)"},
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_elem * n, "avg_launch_ms": ms},
        "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": None, "e2e": None,
        "gpu_reference": ref,
        "scan_variant": rs.get_param("scan_variant") if what == "scan" else None,
    }
    print(json.dumps(line), flush=True)
    return 0


def bit_exact_sharded_check(rs, mgpu, dist, args, rank, world, log2n=24):
    """The sharded sort against the oracle (checker only) at a size the CPU sort finishes in a second:
    concatenated shards == sortByHost of the whole input (Baseline1.cu:15-64)."""
    import torch
    total = (1 << log2n) + 4321
    per = total // world
    first = rank * per
    count = per if rank < world - 1 else total - first
    ok_all = True
    for kind in ("uniform", "zipf"):
        cdf = None
        if kind == "zipf":
            import oracle as O
            cdf = O.zipf_cdf()
        keys = rs.generate(kind, count, first=first, total=total, zipf_cdf=cdf)
        sorter = mgpu.ShardedSorter(dist.group.WORLD, per_rank_capacity=total + 1024, nbits=args.nbits,
                                    fused=not args.no_fused, allow_narrow=not args.no_narrow,
                                    balance_threshold=args.balance_threshold)
        res = sorter.sort(keys)
        sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
        sizes[rank] = res.numel()
        dist.all_reduce(sizes)
        off = int(sizes[:rank].sum().item())
        whole_out = torch.zeros(total, dtype=torch.int32, device="cuda")
        whole_out[off:off + res.numel()] = res
        dist.all_reduce(whole_out)                       # disjoint slices: sum == concatenation
        whole_in = torch.zeros(total, dtype=torch.int32, device="cuda")
        whole_in[first:first + count] = keys
        dist.all_reduce(whole_in)
        ok = torch.ones(1, device="cuda", dtype=torch.int32)
        if rank == 0:
            import oracle as O
            exp = O.sort_keys(whole_in.cpu().numpy().view(np.uint32), args.nbits)
            if not np.array_equal(whole_out.cpu().numpy().view(np.uint32), exp):
                ok.zero_()
        dist.broadcast(ok, 0)
        ok_all = ok_all and bool(ok.item())
        del sorter, res, whole_in, whole_out
    torch.cuda.empty_cache()
    return ok_all


def run_multi(args):
    import torch
    import torch.distributed as dist

    import cuda.radixsort_b200 as rs
    from cuda.radixsort_b200 import mgpu

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rs.load()
    total = 1 << args.log2n_multi
    per = total // world
    pairs = args.workload == "pairs"
    kind = "uniform" if pairs else args.workload
    cdf = None
    if kind == "zipf":
        import oracle as O  # the CDF table of the synthetic Zipf workload (input generation only)
        cdf = O.zipf_cdf()
    keys = rs.generate(kind, per, first=rank * per, total=total, zipf_cdf=cdf)
    vals = torch.arange(rank * per, (rank + 1) * per, dtype=torch.int64, device="cuda").to(torch.int32) if pairs else None
    slack = 1.02 if kind == "uniform" else 1.6      # skewed keys: value splitters never split one value
    bit_exact = None
    if not args.no_bit_exact:
        bit_exact = bit_exact_sharded_check(rs, mgpu, dist, args, rank, world)
    sorter = mgpu.ShardedSorter(dist.group.WORLD, per_rank_capacity=int(per * slack) + (1 << 20),
                                nbits=args.nbits, fused=not args.no_fused, allow_narrow=not args.no_narrow,
                                balance_threshold=args.balance_threshold)

    def sort_once():
        return sorter.sort(keys) if not pairs else sorter.sort_pairs(keys, vals)[0]

    result = None
    for _ in range(args.warmup):
        result = sort_once()
    torch.cuda.synchronize(); dist.barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    launches0 = rs.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    start.record()
    for _ in range(args.steps):
        result = sort_once()
    stop.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([start.elapsed_time(stop)], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([rs.launch_count() - launches0], device="cuda", dtype=torch.int64)
    dist.all_reduce(launches)
    phases = sorter.phase_report()          # per-phase ms, max over ranks
    ok = mgpu.verify_sharded(result, keys, dist.group.WORLD)
    clocks = sampler.stop() if sampler else None

    # e2e: every rank's shard starts and ends in pinned HOST memory (H2D + sharded sort + D2H inside)
    e2e = None
    if not args.no_e2e:
        flag = torch.ones(1, device="cuda", dtype=torch.int32)
        try:
            h_in = torch.empty(per, dtype=torch.int32).pin_memory()
            h_out = torch.empty(int(sorter.capacity), dtype=torch.int32).pin_memory()
            h_in.copy_(keys)
        except Exception:
            flag.zero_()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            d_in = torch.empty_like(keys)

            last = {}

            def e2e_step():
                d_in.copy_(h_in, non_blocking=True)
                res = sorter.sort(d_in)
                h_out[: res.numel()].copy_(res, non_blocking=True)
                torch.cuda.synchronize()
                last["res"] = res
                return res.numel()

            e2e_step()
            e2e_steps = max(1, min(args.steps, 3))
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                got = e2e_step()
            dist.barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            moved = torch.tensor([per * 4, got * 4], device="cuda", dtype=torch.int64)
            dist.all_reduce(moved)
            # what came back to the host is the rank's sorted slice: re-upload it and run the sharded checks
            back = torch.empty(got, dtype=torch.int32, device="cuda")
            back.copy_(h_out[:got])
            e2e_ok = bool(mgpu.verify_sharded(back, keys, dist.group.WORLD))
            del back
            e2e = {"value": total / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(moved[0].item()),
                   "d2h_bytes_per_step": int(moved[1].item()), "ms_per_step": float(dt.item()) * 1e3,
                   "steps": e2e_steps, "verified": e2e_ok,
                   "how": "per rank: pinned host shard -> H2D -> ShardedSorter.sort -> D2H of the rank's sorted "
                          "slice into pinned host memory; wall clock, max over ranks"}
    if rank == 0:
        ms_per_step = float(ms.item()) / args.steps
        line = {
            "metric": METRIC, "value": total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(args, world),
            "phases_ms": phases, "verified": bool(ok), "bit_exact": bit_exact,
            "bit_exact_how": "before the timed region: the same ShardedSorter on 2^24+4321 uniform and Zipf keys, "
                             "concatenated shards compared with the oracle's sortByHost restatement on rank 0",
            "gpu_launches": int(launches.item()),
            "clocks": clocks, "roofline": None, "cpu_baseline": None, "e2e": e2e,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--balance-threshold", type=float, default=1.2,
                    help="multi-GPU: bin-edge splitters leaving a shard above this x mean switch to value splitters")
    ap.add_argument("--workload", default="uniform",
                    choices=["uniform", "zipf", "unique16", "all_equal", "sorted", "reversed", "pairs", "scan", "hist"])
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--log2n-multi", type=int, default=32)
    ap.add_argument("--nbits", type=int, default=8)
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--cpu-sample-log2", type=int, default=26)
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-baselines", action="store_true",
                    help="skip the same-box Parallel7 sortByDevice and cub::DeviceRadixSort timings")
    ap.add_argument("--no-bit-exact", action="store_true", help="multi-GPU: skip the 2^24-key oracle comparison")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-narrow", action="store_true")
    ap.add_argument("--narrow-variant", type=int, default=None)
    ap.add_argument("--scan-variant", type=int, default=None)
    ap.add_argument("--dst-bulk", type=int, default=None, choices=[0, 1],
                    help="multi-GPU fused exchange: 1 = bulk-copy peer stores (default), 0 = 4-byte peer stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)
    if args.variant is not None:
        import cuda.radixsort_b200 as rs
        rs.set_param("variant", args.variant)
    if args.narrow_variant is not None:
        import cuda.radixsort_b200 as rs
        rs.set_param("narrow_variant", args.narrow_variant)
    if args.dst_bulk is not None:
        import cuda.radixsort_b200 as rs
        rs.set_param("dst_bulk", args.dst_bulk)
    if args.scan_variant is not None:
        import cuda.radixsort_b200 as rs
        rs.set_param("scan_variant", args.scan_variant)
    if args.workload in ("scan", "hist"):
        return run_primitive(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        if world == 1:
            print(json.dumps({"error": "--gpus N > 1 must be launched with torch.distributed.run"}))
            return 2
        return run_multi(args)
    return run_single(args)


if __name__ == "__main__":
    sys.exit(main())
