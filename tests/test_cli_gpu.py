"""GPU: the reference-style command line driver (csrc/radixsort_cli.cpp) -- the reference's whole
test procedure (SourceCode/Parallel7.cu:696-775: sort by host, sort by device, compare) running
against the new library through include/radix_sort_compat.hpp."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda", "radixsort_b200", "radixsort_cli")


def run_cli(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([CLI] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=300)


def test_cli_log_format_and_exit_code(rs):
    r = run_cli([512, 8, 513])
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for needle in ("**********GPU info**********", "Name: ", "Num SMs: ", "Input size: 513", "Block size: 512",
                   "Digit width: 8-bit", "Radix Sort by host", "Radix Sort by device:", "Time: "):
        assert needle in out, needle
    assert out.count("CORRECT :)") == 2 and "INCORRECT" not in out


@pytest.mark.parametrize("args", [(512, 8), (256, 4, 100001), (1024, 5, 7681), (512, 11, 40000)])
def test_cli_against_the_compiled_reference(rs, oracle, args):
    from oracle import build_oracle
    ref = build_oracle.ref_so("Baseline1")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/libref_baseline1.so not present")
    r = run_cli(args, {"B200SORT_REF_LIB": ref})
    assert r.returncode == 0, r.stdout + r.stderr
    assert ref in r.stdout and r.stdout.count("CORRECT :)") == 2


def test_cli_rejects_bad_digit_width_like_check(rs):
    r = run_cli([512, 0, 100])
    assert r.returncode != 0
    assert "Error:" in r.stderr and "code:" in r.stderr and "reason:" in r.stderr    # CHECK format, common.h:6-16


def test_cli_sharded_over_all_gpus(rs):
    # B200SORT_GPUS: sortByDevice goes through b200sort_mgpu_keys_host (every visible device)
    r = run_cli([512, 8, (1 << 22) + 3], {"B200SORT_GPUS": "0"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GPUs: all" in r.stdout and r.stdout.count("CORRECT :)") == 2
