"""ctypes binding of libb200sort.so -- the only compute path of this package.

There is no fallback: if the library is missing or cannot be loaded, every entry point raises
``RadixSortUnavailable``.  Signatures follow include/b200sort.h one to one.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libb200sort.so")

U32P = C.POINTER(C.c_uint32)
U64P = C.POINTER(C.c_uint64)


class RadixSortError(RuntimeError):
    """A libb200sort call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"b200sort error {code}: {message}")
        self.code = code


class RadixSortUnavailable(RuntimeError):
    """libb200sort.so is not built / not loadable (no CPU fallback exists)."""


# name -> (restype, argtypes); every symbol include/b200sort.h declares.
SIGNATURES = {
    "b200sort_keys_host": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]),
    "b200sort_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int]),
    "b200sort_shutdown": (C.c_int, []),
    "b200sort_warmup": (C.c_int, [C.c_uint64, C.c_int]),
    "b200sort_route": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "b200sort_mgpu_keys_host": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int,
                                          C.POINTER(C.c_int), C.c_int]),
    "b200sort_mgpu_pairs_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "b200sort_mgpu_last_stats": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "b200sort_mgpu_shutdown": (C.c_int, []),
    "b200sort_plan_owners": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "b200sort_plan_value_cuts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "b200sort_temp_bytes": (C.c_size_t, [C.c_uint64, C.c_int, C.c_int]),
    "b200sort_keys": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                C.c_void_p]),
    "b200sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "b200sort_keys_low_bits": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                         C.c_void_p]),
    "b200sort_pairs_low_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    "b200sort_histogram": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    "b200sort_digit_pass": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "b200sort_scan_temp_bytes": (C.c_size_t, [C.c_uint64]),
    "b200sort_exclusive_scan": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200sort_generate": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64,
                                    C.c_void_p, C.c_void_p]),
    "b200sort_store_probe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]),
    "b200sort_verify": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "b200sort_profile_enable": (C.c_int, [C.c_int]),
    "b200sort_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "b200sort_launch_count": (C.c_uint64, []),
    "b200sort_set_param": (C.c_int, [C.c_char_p, C.c_int]),
    "b200sort_get_param": (C.c_int, [C.c_char_p]),
    "b200sort_tile_keys": (C.c_int, [C.c_int]),
    "b200sort_algorithmic_bytes": (C.c_uint64, [C.c_uint64, C.c_int, C.c_int]),
    "b200sort_num_passes": (C.c_int, [C.c_int]),
    "b200sort_device_banner": (C.c_int, [C.c_char_p, C.c_size_t]),
    "b200sort_version": (C.c_int, []),
    "b200sort_error_string": (C.c_char_p, [C.c_int]),
    "b200sort_last_error_string": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libb200sort.so (building it first is the job of __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadixSortUnavailable(
            f"{LIB_PATH} not found: run `python -m cuda.radixsort_b200.build` "
            "(there is no CPU fallback)")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover - depends on the host
        raise RadixSortUnavailable(f"cannot load {LIB_PATH}: {e}") from e
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        lib = load()
        msg = lib.b200sort_last_error_string().decode() or lib.b200sort_error_string(code).decode()
        raise RadixSortError(code, msg)
