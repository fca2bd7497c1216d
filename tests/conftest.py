"""Test configuration.

  -m "not gpu"  oracle vs the reference's known answers / golden fixtures, host logic, C-ABI
                symbol and error-path checks (no compute call), world_size-2 gloo tests.
  -m gpu        parity tests proper: the CUDA path, called through the C ABI, against the
                oracle on the same inputs (bit-exact: this is integer work).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    return O


@pytest.fixture(scope="session")
def rs():
    """The product package with the library built and loaded (fails loudly if it cannot be)."""
    import __graft_entry__ as entry
    entry.build(reference=False)
    import cuda.radixsort_b200 as rs_mod
    rs_mod.load()
    return rs_mod


def to_dev(a: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def to_host(t) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)
