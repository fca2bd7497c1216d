"""In-tree build of libb200sort.so (hand-written sm_100a CUDA behind a C ABI).

    python -m cuda.radixsort_b200.build [--force] [--verbose]

One nvcc invocation per translation unit (the digit-pass kernels are compiled once per digit
width so the eight widths build in parallel), then one link step.  Objects go to
cuda/radixsort_b200/_build/, the library to cuda/radixsort_b200/libb200sort.so -- inside the
tree so it travels to the GPU box with the repo snapshot.  The CUDA runtime is linked
statically: the library has no dependency beyond libstdc++ and the driver.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "_build")
LIB_PATH = os.path.join(PKG_DIR, "libb200sort.so")
CLI_PATH = os.path.join(PKG_DIR, "radixsort_cli")
INCLUDE_DIR = os.path.normpath(os.path.join(PKG_DIR, "..", "..", "include"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-pthread", "-I", INCLUDE_DIR]
if os.environ.get("B200_TUNING"):
    NVCC_FLAGS.append("-DB200_TUNING")  # every kernel geometry of csrc/launch.h (tools/sweep.py, the variant tests)
if os.environ.get("B200_COL_NOTMA"):
    NVCC_FLAGS.append("-DB200_COL_NOTMA")  # experiment: tile loaded with LDG + STS instead of the bulk-copy engine
if os.environ.get("B200_COL_DEBUG"):
    NVCC_FLAGS.append("-DB200_COL_DEBUG")  # per-phase clock stamps in the column-sweep kernel (tools/col_timeline.py)
WIDTHS = range(1, 9)


def _nvcc() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _sources_mtime() -> float:
    newest = 0.0
    for d in (CSRC, INCLUDE_DIR):
        for name in os.listdir(d):
            newest = max(newest, os.path.getmtime(os.path.join(d, name)))
    return max(newest, os.path.getmtime(os.path.abspath(__file__)))


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"command failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout, r.stderr, flush=True)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    """Compile (if stale) and return the path of libb200sort.so."""
    flags_file = os.path.join(BUILD_DIR, "flags.txt")
    flags_now = " ".join(NVCC_FLAGS)
    try:
        same_flags = open(flags_file).read() == flags_now
    except OSError:
        same_flags = False
    if not force and same_flags and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _sources_mtime():
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    jobs = []
    objs = []
    for w in WIDTHS:
        obj = os.path.join(BUILD_DIR, f"kernels_w{w}.o")
        objs.append(obj)
        jobs.append([nvcc, *NVCC_FLAGS, *extra, f"-DB200_W={w}", "-c",
                     os.path.join(CSRC, "kernels_w.cu"), "-o", obj])
    obj = os.path.join(BUILD_DIR, "b200sort.o")
    objs.append(obj)
    jobs.append([nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, "b200sort.cu"), "-o", obj])
    obj = os.path.join(BUILD_DIR, "mgpu_host.o")
    objs.append(obj)
    jobs.append([nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "mgpu_host.cu"), "-o", obj])
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        list(pool.map(lambda c: _run(c, verbose or ptxas_info), jobs))
    _run([nvcc, *ARCH, "-shared", "-Xcompiler", "-pthread", "-o", LIB_PATH, *objs], verbose)
    with open(flags_file, "w") as f:
        f.write(flags_now)
    build_cli(verbose)
    return LIB_PATH


def build_cli(verbose: bool = False) -> str | None:
    """The reference-style command line driver (C++ host code above the C ABI)."""
    src = os.path.join(CSRC, "radixsort_cli.cpp")
    if not os.path.exists(src):
        return None
    cmd = ["g++", "-O2", "-std=c++17", "-I", INCLUDE_DIR, src, "-o", CLI_PATH,
           f"-L{PKG_DIR}", "-lb200sort", f"-Wl,-rpath,{PKG_DIR}", "-Wl,-rpath,$ORIGIN", "-ldl"]
    _run(cmd, verbose)
    return CLI_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                 ptxas_info="--ptxas" in sys.argv)
    print(path)
