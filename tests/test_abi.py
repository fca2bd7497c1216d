"""CPU: the C-ABI library loads, exports every symbol include/b200sort.h declares, and its
argument checking / bookkeeping entry points behave (no compute call without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200sort.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sort_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("b200sort_keys_host", "b200sort_pairs_host", "b200sort_keys", "b200sort_pairs",
                 "b200sort_temp_bytes", "b200sort_histogram", "b200sort_digit_pass",
                 "b200sort_mgpu_keys_host", "b200sort_mgpu_pairs_host"):
        assert must in syms


def test_library_exports_every_declared_symbol(rs):
    lib = C.CDLL(rs.LIB_PATH)
    from cuda.radixsort_b200 import _lib
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in b200sort.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_oracle_or_library_sort_linked(rs):
    # the product must not carry a CPU sort or CUB/Thrust: look at the dynamic symbol table
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", rs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out
    full = subprocess.run(["nm", "-C", rs.LIB_PATH], capture_output=True, text=True).stdout
    assert "cub::" not in full and "thrust::" not in full


def test_pass_counts(rs):
    # widths 1..8: one kernel pass per reference digit; 9..16: two sub-digit passes per digit
    for nbits in range(1, 9):
        assert rs.num_passes(nbits) == -(-32 // nbits)
    assert rs.num_passes(9) == 7      # digits 9,9,9,5 -> (5,4) x3 + (5)
    assert rs.num_passes(12) == 6     # 12,12,8 -> (6,6),(6,6),(6,2)
    assert rs.num_passes(16) == 4
    for bad in (0, -1, 17, 33):
        with pytest.raises(rs.RadixSortError):
            rs.num_passes(bad)


def test_algorithmic_bytes_formula(rs):
    n = 1 << 28
    assert rs.algorithmic_bytes(n, 8, False) == 4 * n * 9 == 9663676416
    assert rs.algorithmic_bytes(n, 8, True) == 4 * n * 17 == 18253611008
    assert rs.algorithmic_bytes(n, 4, False) == 4 * n * 17


def test_temp_bytes_monotone_and_covers_alt_buffers(rs):
    prev = 0
    for n in (0, 1, 1000, 1 << 20, 1 << 28):
        t = rs.temp_bytes(n, 8, False)
        assert t >= prev and t >= 4 * n
        assert rs.temp_bytes(n, 8, True) >= t + 4 * n
        prev = t
    assert rs.temp_bytes(1 << 28, 8, False) < 1.3 * 4 * (1 << 28)


def test_argument_errors_without_touching_the_gpu(rs):
    lib = rs.load()
    a = np.zeros(16, np.uint32)
    o = np.zeros(16, np.uint32)
    assert lib.b200sort_keys_host(a.ctypes.data, 16, o.ctypes.data, 0, 512) == -1      # nBits
    assert lib.b200sort_keys_host(a.ctypes.data, 16, o.ctypes.data, 17, 512) == -1
    assert lib.b200sort_keys_host(a.ctypes.data, 16, o.ctypes.data, 8, 0) == -1        # blockSize
    assert lib.b200sort_keys_host(a.ctypes.data, 1 << 32, o.ctypes.data, 8, 512) == -2  # too big
    assert lib.b200sort_keys_host(None, 16, o.ctypes.data, 8, 512) == -1
    assert lib.b200sort_keys_host(a.ctypes.data, 0, o.ctypes.data, 8, 512) == 0        # n == 0 no-op
    assert lib.b200sort_keys(None, 0, None, None, 0, 8, None) == 0
    assert lib.b200sort_keys_low_bits(None, 0, None, None, 0, 8, 29, None) == 0
    assert lib.b200sort_keys_low_bits(None, 0, None, None, 0, 8, 0, None) == -1       # key_bits
    assert lib.b200sort_pairs_low_bits(None, None, 0, None, None, None, 0, 8, 33, None) == -1
    # multi-GPU host entry points validate before they look for devices
    assert lib.b200sort_mgpu_keys_host(a.ctypes.data, 16, o.ctypes.data, 0, 512, None, 2) == -1
    assert lib.b200sort_mgpu_keys_host(a.ctypes.data, 16, o.ctypes.data, 8, 0, None, 2) == -1
    assert lib.b200sort_mgpu_last_stats(None, 0) == 0
    assert lib.b200sort_mgpu_shutdown() == 0
    assert lib.b200sort_set_param(b"variant", 9999) == -1
    assert lib.b200sort_set_param(b"variant", -1) == 0     # automatic choice
    # L2 prefetch distances of the digit pass and of the scan: -1 = automatic, 0 = off
    for name in (b"prefetch_tiles", b"scan_prefetch_tiles"):
        assert lib.b200sort_get_param(name) == -1
        assert lib.b200sort_set_param(name, 0) == 0 and lib.b200sort_get_param(name) == 0
        assert lib.b200sort_set_param(name, 64) == 0 and lib.b200sort_get_param(name) == 64
        assert lib.b200sort_set_param(name, -2) == -1
        assert lib.b200sort_set_param(name, -1) == 0
    assert lib.b200sort_set_param(b"no_such_param", 1) == -1
    assert b"invalid" in lib.b200sort_error_string(-1)
    assert lib.b200sort_version() == 100


def test_reference_interface_mirror(rs):
    # same names / argument order as SourceCode/Parallel7.cu:22,530,641-645
    assert [m.name for m in rs.Implementation] == ["SORT_BY_HOST", "SORT_BY_THRUST", "SORT_BY_DEVICE"]
    assert int(rs.SORT_BY_DEVICE) == 2
    import inspect
    assert list(inspect.signature(rs.sort).parameters) == ["in_", "n", "out", "implementation", "numBits", "blockSize"]
    assert inspect.signature(rs.sort).parameters["numBits"].default == 4
    assert inspect.signature(rs.sort).parameters["blockSize"].default == 1
    assert list(inspect.signature(rs.sortByDevice).parameters) == ["h_input", "n", "h_output", "numBits", "blockSize"]
    a = np.zeros(4, np.uint32)
    for impl in (rs.SORT_BY_HOST, rs.SORT_BY_THRUST, False):
        with pytest.raises(rs.RadixSortError):   # no CPU path in the product
            rs.sort(a, 4, a.copy(), impl, 8)
    with pytest.raises(TypeError):
        rs.sortByDevice(np.zeros(4, np.int64), 4, a, 8, 512)


def test_product_does_not_import_the_oracle():
    import ast
    pkg = os.path.join(ROOT, "cuda", "radixsort_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            tree = ast.parse(open(os.path.join(pkg, name)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), name
    for name in os.listdir(os.path.join(pkg, "csrc")):
        src = open(os.path.join(pkg, "csrc", name)).read()
        code = re.sub(r"//.*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))   # comments may cite it
        assert "oracle" not in code, name


def test_headers_compile_as_plain_c_and_cxx():
    """include/b200sort.h is a C header (the boundary a C, Go/cgo or JNI host would bind);
    include/radix_sort_compat.hpp is the C++ shim with the reference's spellings."""
    import shutil
    import subprocess
    inc = os.path.dirname(HEADER)
    if shutil.which("gcc"):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    if shutil.which("g++"):
        r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++",
                            os.path.join(inc, "radix_sort_compat.hpp")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
