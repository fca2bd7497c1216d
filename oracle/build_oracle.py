"""Build recipe for the test oracle (TEST INFRASTRUCTURE, not product code).

  oracle/liboracle.so              gcc build of oracle/radix_oracle.c (always)
  oracle/_ref/libref_<file>.so     the UNMODIFIED reference translation units
                                   /root/reference/SourceCode/{Baseline1,Baseline4,Parallel7}.cu
                                   compiled where they lie (never copied into the repo), one
                                   shared object per file because every file defines the same
                                   global symbols.  Only built when /root/reference exists (the
                                   build container); the GPU box receives the prebuilt files.

The reference has no build system; the implied build is `nvcc -I SourceCode File.cu`
(SURVEY.md section 0).  `-Dmain=ref_main` keeps its main() out of the way.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("B200SORT_REFERENCE", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_FILES = ("Baseline1", "Baseline4", "Parallel7")
# the reference's standalone kernel studies (Docs/Snippets), used only as same-box bars for the scan and
# histogram primitives in bench.py --workload scan|hist
SNIPPET_FILES = {"PrefixSum": "PrefixSum-WorkEfficient", "Histogram": "Histogram"}


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "radix_oracle.c")
    if force or not _newer(ORACLE_SO, src):
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
               "-o", ORACLE_SO, src]
        subprocess.run(cmd, check=True)
    return ORACLE_SO


def ref_so(name: str) -> str:
    return os.path.join(REF_DIR, f"libref_{name.lower()}.so")


def build_reference(force: bool = False, files=REF_FILES + tuple(SNIPPET_FILES)) -> list[str]:
    """Compile the reference's own .cu files into oracle/_ref/ (skipped without the sources)."""
    src_dir = os.path.join(REF_ROOT, "SourceCode")
    if not os.path.isdir(src_dir):
        return [ref_so(f) for f in files if os.path.exists(ref_so(f))]
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(REF_DIR, exist_ok=True)
    built = []
    for name in files:
        if name in SNIPPET_FILES:
            inc = os.path.join(REF_ROOT, "Docs", "Snippets")
            src = os.path.join(inc, f"{SNIPPET_FILES[name]}.cu")
        else:
            inc = src_dir
            src = os.path.join(src_dir, f"{name}.cu")
        out = ref_so(name)
        if force or not _newer(out, src):
            cmd = [nvcc, "-O2", "-std=c++17", "-w",
                   "-gencode", "arch=compute_100a,code=sm_100a",
                   "-I", inc, "-Dmain=ref_main",
                   "-Xcompiler", "-fPIC", "-shared", "-o", out, src]
            subprocess.run(cmd, check=True)
        built.append(out)
    return built


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_oracle(force))
    if "--no-ref" not in sys.argv:
        for p in build_reference(force):
            print(p)
