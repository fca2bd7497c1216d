"""cuda.radixsort_b200 -- B200 (sm_100a) LSD radix sort behind the interface of
truongchauhien/CUDA.RadixSort's device sort path.

The compute path is libb200sort.so (hand-written CUDA behind the C ABI of include/b200sort.h).
This package is the thin host-side mirror of the reference interface plus the multi-GPU
driver; importing it never touches the GPU, and calling it without the built library raises
(there is no CPU fallback).  `cuda` is a namespace package shared with cuda-python.
"""
from ._lib import LIB_PATH, RadixSortError, RadixSortUnavailable, load  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import DEFAULT_NBITS  # noqa: F401
