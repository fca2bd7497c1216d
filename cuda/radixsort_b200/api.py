"""Host-side mirror of the reference's sort interface, on top of the C ABI.

Reference interface (SourceCode/Parallel7.cu:22, :530, :641-645)::

    typedef enum { SORT_BY_HOST, SORT_BY_THRUST, SORT_BY_DEVICE } Implementation;
    void sortByDevice(const uint32_t *h_input, int n, uint32_t *h_output, int numBits, int blockSize);
    void sort(const uint32_t *in, int n, uint32_t *out,
              Implementation implementation = SORT_BY_HOST, int numBits = 4, int blockSize = 1);

Same names, argument order and meaning here.  Differences, all deliberate:
  * SORT_BY_HOST / SORT_BY_THRUST raise: this package has no CPU path and links no Thrust
    (the reference's sortByHost lives on as the test oracle under oracle/, outside the product);
  * errors raise RadixSortError instead of print-and-exit (common/common.h:6-16);
  * n may exceed 2^31-1 (up to 2^32-1).

PyTorch is used for device memory and streams only (torch.Tensor.data_ptr / current stream).
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import RadixSortError

__all__ = [
    "Implementation", "SORT_BY_HOST", "SORT_BY_THRUST", "SORT_BY_DEVICE", "sort", "sortByDevice",
    "sort_by_device", "sort_pairs_by_device", "sort_by_devices", "sort_pairs_by_devices", "mgpu_last_stats", "mgpu_shutdown", "warmup",
    "Workspace", "sort_keys", "sort_pairs", "histogram",
    "digit_pass", "route", "exclusive_scan", "generate", "verify", "temp_bytes", "algorithmic_bytes", "num_passes",
    "tile_keys", "set_param", "get_param", "profile_enable", "profile_read", "launch_count",
    "shutdown",
]


class Implementation(enum.IntEnum):
    SORT_BY_HOST = 0
    SORT_BY_THRUST = 1
    SORT_BY_DEVICE = 2


SORT_BY_HOST = Implementation.SORT_BY_HOST
SORT_BY_THRUST = Implementation.SORT_BY_THRUST
SORT_BY_DEVICE = Implementation.SORT_BY_DEVICE

GEN_KINDS = {"uniform": 0, "zipf": 1, "unique16": 2, "all_equal": 3, "sorted": 4, "reversed": 5,
             "iota": 6}


# ---------------------------------------------------------------------------------------------
# host-pointer path (numpy arrays stand in for the reference's malloc'ed uint32_t*)

def _host_u32(a, name: str, writable: bool = False) -> np.ndarray:
    if not isinstance(a, np.ndarray) or a.dtype != np.uint32 or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a C-contiguous numpy uint32 array")
    if writable and not a.flags.writeable:
        raise TypeError(f"{name} must be writable")
    return a


def sortByDevice(h_input: np.ndarray, n: int, h_output: np.ndarray, numBits: int, blockSize: int) -> None:
    """Mirror of sortByDevice (SourceCode/Parallel7.cu:530): host arrays in, host arrays out."""
    h_input = _host_u32(h_input, "h_input")
    h_output = _host_u32(h_output, "h_output", writable=True)
    if n < 0 or n > h_input.size or n > h_output.size:
        raise ValueError("n exceeds the arrays")
    lib = _lib.load()
    _lib.check(lib.b200sort_keys_host(h_input.ctypes.data, n, h_output.ctypes.data, numBits, blockSize))


sort_by_device = sortByDevice


def sort_pairs_by_device(h_keys_in, h_vals_in, n, h_keys_out, h_vals_out, numBits: int, blockSize: int) -> None:
    """Stable key/value variant of sortByDevice (north-star extension; no reference counterpart)."""
    ki, vi = _host_u32(h_keys_in, "h_keys_in"), _host_u32(h_vals_in, "h_vals_in")
    ko, vo = _host_u32(h_keys_out, "h_keys_out", True), _host_u32(h_vals_out, "h_vals_out", True)
    if n < 0 or n > min(ki.size, vi.size, ko.size, vo.size):
        raise ValueError("n exceeds the arrays")
    lib = _lib.load()
    _lib.check(lib.b200sort_pairs_host(ki.ctypes.data, vi.ctypes.data, n, ko.ctypes.data, vo.ctypes.data,
                                       numBits, blockSize))


def _device_list(devices):
    import ctypes as C
    if devices is None:
        return None, 0
    if isinstance(devices, int):
        return None, int(devices)
    devices = [int(d) for d in devices]
    return (C.c_int * len(devices))(*devices), len(devices)


def sort_by_devices(h_input, n: int, h_output, numBits: int, blockSize: int, devices=None) -> None:
    """sortByDevice over several GPUs of this node from one process (b200sort_mgpu_keys_host).

    ``devices``: list of CUDA ordinals (repeats allowed: the shards then share that GPU), an
    int (the first that many devices) or None (every visible device)."""
    h_input = _host_u32(h_input, "h_input")
    h_output = _host_u32(h_output, "h_output", writable=True)
    if n < 0 or n > h_input.size or n > h_output.size:
        raise ValueError("n exceeds the arrays")
    arr, count = _device_list(devices)
    lib = _lib.load()
    _lib.check(lib.b200sort_mgpu_keys_host(h_input.ctypes.data, n, h_output.ctypes.data, numBits, blockSize,
                                           arr, count))


def sort_pairs_by_devices(h_keys_in, h_vals_in, n, h_keys_out, h_vals_out, numBits: int, blockSize: int,
                          devices=None) -> None:
    """Stable key/value form of :func:`sort_by_devices` (b200sort_mgpu_pairs_host)."""
    ki, vi = _host_u32(h_keys_in, "h_keys_in"), _host_u32(h_vals_in, "h_vals_in")
    ko, vo = _host_u32(h_keys_out, "h_keys_out", True), _host_u32(h_vals_out, "h_vals_out", True)
    if n < 0 or n > min(ki.size, vi.size, ko.size, vo.size):
        raise ValueError("n exceeds the arrays")
    arr, count = _device_list(devices)
    lib = _lib.load()
    _lib.check(lib.b200sort_mgpu_pairs_host(ki.ctypes.data, vi.ctypes.data, n, ko.ctypes.data, vo.ctypes.data,
                                            numBits, blockSize, arr, count))


MGPU_STAT_NAMES = ("upload_ms", "histogram_ms", "plan_ms", "partition_ms", "exchange_wait_ms", "local_sort_ms",
                   "download_ms", "partition_shift", "partition_bits", "imbalance", "devices", "value_splitters")


def mgpu_last_stats() -> dict:
    """Device-event figures of the last sort_by_devices / sort_pairs_by_devices call."""
    import ctypes as C
    buf = (C.c_double * len(MGPU_STAT_NAMES))()
    k = _lib.load().b200sort_mgpu_last_stats(buf, len(MGPU_STAT_NAMES))
    return {name: buf[i] for i, name in enumerate(MGPU_STAT_NAMES[:k])}


def sort(in_, n: int, out, implementation=SORT_BY_HOST, numBits: int = 4, blockSize: int = 1) -> None:
    """Mirror of sort() (SourceCode/Parallel7.cu:641-662).

    ``implementation`` may also be a bool, the north star's ``useDevice`` spelling
    (``sort(in, n, out, useDevice, blockSize)``); then the fifth positional argument is the
    block size and the digit width is the module default ``DEFAULT_NBITS`` (8).
    """
    if isinstance(implementation, (bool, np.bool_)):
        block = numBits  # fifth positional argument of the bool spelling is the block size
        if not implementation:
            raise RadixSortError(-1, "useDevice=false: this build has no CPU path "
                                     "(see oracle/ for the test oracle)")
        return sortByDevice(in_, n, out, DEFAULT_NBITS, block)
    impl = Implementation(implementation)
    if impl != SORT_BY_DEVICE:
        raise RadixSortError(-1, f"{impl.name}: this build has no CPU / Thrust path (see oracle/ for the test oracle)")
    return sortByDevice(in_, n, out, numBits, blockSize)


DEFAULT_NBITS = 8


# ---------------------------------------------------------------------------------------------
# device-resident path (torch tensors carry device memory; any 4-byte integer dtype)

def _torch():
    import torch
    return torch


def _dev_ptr(t, name: str) -> int:
    torch = _torch()
    if not isinstance(t, torch.Tensor) or not t.is_cuda or not t.is_contiguous() or t.element_size() != 4:
        raise TypeError(f"{name} must be a contiguous CUDA tensor of a 4-byte dtype")
    return t.data_ptr()


def _stream_ptr(stream) -> int:
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def temp_bytes(n: int, nbits: int = 8, pairs: bool = False) -> int:
    return int(_lib.load().b200sort_temp_bytes(n, nbits, int(pairs)))


def algorithmic_bytes(n: int, nbits: int = 8, pairs: bool = False) -> int:
    return int(_lib.load().b200sort_algorithmic_bytes(n, nbits, int(pairs)))


def num_passes(nbits: int) -> int:
    r = _lib.load().b200sort_num_passes(nbits)
    if r < 0:
        raise RadixSortError(r, "nBits must be in 1..16")
    return r


def tile_keys(pairs: bool = False) -> int:
    return _lib.load().b200sort_tile_keys(int(pairs))


class Workspace:
    """Caller-owned temp storage for the device-resident entry points (grows on demand)."""

    def __init__(self, device=None):
        self.device = device
        self.buf = None

    def get(self, nbytes: int):
        torch = _torch()
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device or "cuda")
        assert self.buf.data_ptr() % 256 == 0
        return self.buf


_default_ws: dict = {}


def _workspace(device) -> Workspace:
    key = str(device)
    if key not in _default_ws:
        _default_ws[key] = Workspace(device)
    return _default_ws[key]


def sort_keys(keys, nbits: int = 8, out=None, workspace: Workspace | None = None, stream=None, key_bits: int = 32):
    """Device-resident sort: b200sort_keys.  Asynchronous on the current (or given) stream.
    key_bits < 32: the caller knows that all keys agree in their bits >= key_bits (b200sort_keys_low_bits)."""
    torch = _torch()
    kp = _dev_ptr(keys, "keys")
    n = keys.numel()
    if out is None:
        out = torch.empty_like(keys)
    op = _dev_ptr(out, "out")
    ws = workspace or _workspace(keys.device)
    need = temp_bytes(n, nbits, False)
    tmp = ws.get(need)
    lib = _lib.load()
    if key_bits == 32:
        _lib.check(lib.b200sort_keys(kp, n, op, tmp.data_ptr(), tmp.numel(), nbits, _stream_ptr(stream)))
    else:
        _lib.check(lib.b200sort_keys_low_bits(kp, n, op, tmp.data_ptr(), tmp.numel(), nbits, key_bits,
                                              _stream_ptr(stream)))
    return out


def sort_pairs(keys, vals, nbits: int = 8, out_keys=None, out_vals=None,
               workspace: Workspace | None = None, stream=None, key_bits: int = 32):
    """Device-resident stable key/value sort: b200sort_pairs (key_bits < 32: b200sort_pairs_low_bits)."""
    torch = _torch()
    kp, vp = _dev_ptr(keys, "keys"), _dev_ptr(vals, "vals")
    n = keys.numel()
    if vals.numel() != n:
        raise ValueError("keys and vals differ in length")
    out_keys = torch.empty_like(keys) if out_keys is None else out_keys
    out_vals = torch.empty_like(vals) if out_vals is None else out_vals
    ws = workspace or _workspace(keys.device)
    tmp = ws.get(temp_bytes(n, nbits, True))
    lib = _lib.load()
    if key_bits == 32:
        _lib.check(lib.b200sort_pairs(kp, vp, n, _dev_ptr(out_keys, "out_keys"), _dev_ptr(out_vals, "out_vals"),
                                      tmp.data_ptr(), tmp.numel(), nbits, _stream_ptr(stream)))
    else:
        _lib.check(lib.b200sort_pairs_low_bits(kp, vp, n, _dev_ptr(out_keys, "out_keys"), _dev_ptr(out_vals, "out_vals"),
                                               tmp.data_ptr(), tmp.numel(), nbits, key_bits, _stream_ptr(stream)))
    return out_keys, out_vals


def histogram(keys, shift: int, bits: int, workspace: Workspace | None = None, stream=None):
    """Digit histogram (uint32 counts as an int32 tensor of 2^bits entries): b200sort_histogram."""
    torch = _torch()
    hist = torch.empty(1 << bits, dtype=torch.int32, device=keys.device)
    ws = workspace or _workspace(keys.device)
    tmp = ws.get(max(4096, temp_bytes(0, bits, False)))
    lib = _lib.load()
    _lib.check(lib.b200sort_histogram(_dev_ptr(keys, "keys"), keys.numel(), shift, bits, hist.data_ptr(),
                                      tmp.data_ptr(), tmp.numel(), _stream_ptr(stream)))
    return hist


def digit_pass(keys, shift: int, bits: int, vals=None, out_keys=None, out_vals=None, bin_dst=None,
               workspace: Workspace | None = None, stream=None):
    """One stable digit pass / MSD partition: b200sort_digit_pass."""
    torch = _torch()
    n = keys.numel()
    lib = _lib.load()
    ws = workspace or _workspace(keys.device)
    tmp = ws.get(temp_bytes(n, bits, vals is not None))
    if bin_dst is None:
        out_keys = torch.empty_like(keys) if out_keys is None else out_keys
        if vals is not None:
            out_vals = torch.empty_like(vals) if out_vals is None else out_vals
    okp = _dev_ptr(out_keys, "out_keys") if out_keys is not None else None
    ovp = _dev_ptr(out_vals, "out_vals") if out_vals is not None else None
    dstp = None
    if bin_dst is not None:
        if bin_dst.dtype != torch.int64 or not bin_dst.is_cuda or bin_dst.numel() < (2 if vals is not None else 1) << bits:
            raise TypeError("bin_dst must be a CUDA int64 tensor of 2^bits (keys) or 2*2^bits (pairs) addresses")
        dstp = bin_dst.data_ptr()
    _lib.check(lib.b200sort_digit_pass(_dev_ptr(keys, "keys"), _dev_ptr(vals, "vals") if vals is not None else None,
                                       n, okp, ovp, shift, bits, dstp, tmp.data_ptr(), tmp.numel(),
                                       _stream_ptr(stream)))
    return (out_keys, out_vals) if vals is not None else out_keys


def route(keys, values, ties=None, out=None, stream=None, with_counts: bool = False):
    """out[i] = #{j : (values[j], ties[j]) <= (keys[i], i)} (unsigned keys, lexicographic): b200sort_route.

    ``values``: non-decreasing integers in [0, 2^32]; ``ties``: for each cut the local index from
    which a key EQUAL to the value counts as at-or-above the cut (default 0: the whole run).
    ``with_counts``: also return the number of keys per destination (int32 tensor, len(values)+1)."""
    torch = _torch()
    values = [int(v) for v in values]
    ties = [0] * len(values) if ties is None else [int(t) for t in ties]
    if len(ties) != len(values):
        raise ValueError("values and ties differ in length")
    table = torch.tensor(values + ties, dtype=torch.int64, device=keys.device)
    out = torch.empty_like(keys) if out is None else out
    counts = torch.empty(len(values) + 1, dtype=torch.int32, device=keys.device) if with_counts else None
    _lib.check(_lib.load().b200sort_route(_dev_ptr(keys, "keys"), keys.numel(),
                                          table.data_ptr() if table.numel() else None,
                                          len(values), _dev_ptr(out, "out"),
                                          counts.data_ptr() if with_counts else None, _stream_ptr(stream)))
    return (out, counts) if with_counts else out


def exclusive_scan(x, out=None, workspace: Workspace | None = None, stream=None):
    """Device-wide exclusive prefix sum of a 4-byte integer tensor (mod 2^32): b200sort_exclusive_scan."""
    torch = _torch()
    n = x.numel()
    out = torch.empty_like(x) if out is None else out
    ws = workspace or _workspace(x.device)
    lib = _lib.load()
    tmp = ws.get(int(lib.b200sort_scan_temp_bytes(n)))
    _lib.check(lib.b200sort_exclusive_scan(_dev_ptr(x, "x"), n, _dev_ptr(out, "out"), tmp.data_ptr(), tmp.numel(),
                                           _stream_ptr(stream)))
    return out


_zipf_cdf_dev: dict = {}


def generate(kind: str, count: int, first: int = 0, total: int | None = None, device=None, out=None,
             zipf_cdf=None, stream=None):
    """Synthetic workloads of SURVEY.md 8d, generated on the device: b200sort_generate.

    ``zipf_cdf``: 65536-entry uint32 table (numpy or CUDA tensor); required for kind="zipf"
    (the table is produced by the test oracle so host and device agree bit for bit).
    """
    torch = _torch()
    device = device or "cuda"
    if out is None:
        out = torch.empty(count, dtype=torch.int32, device=device)
    cdf_ptr = None
    if kind == "zipf":
        if zipf_cdf is None:
            raise ValueError("kind='zipf' needs zipf_cdf")
        if not isinstance(zipf_cdf, torch.Tensor):
            key = (str(out.device), zipf_cdf.ctypes.data)
            if key not in _zipf_cdf_dev:
                _zipf_cdf_dev[key] = torch.from_numpy(zipf_cdf.view(np.int32)).to(out.device)
            zipf_cdf = _zipf_cdf_dev[key]
        cdf_ptr = zipf_cdf.data_ptr()
    lib = _lib.load()
    _lib.check(lib.b200sort_generate(_dev_ptr(out, "out"), first, count, GEN_KINDS[kind],
                                     total if total is not None else first + count, cdf_ptr,
                                     _stream_ptr(stream)))
    return out


def verify(keys, stream=None):
    """(inversions, sum key, sum sm64(key), xor sm64(key)) of a device array: b200sort_verify."""
    torch = _torch()
    res = torch.empty(4, dtype=torch.int64, device=keys.device)
    lib = _lib.load()
    _lib.check(lib.b200sort_verify(_dev_ptr(keys, "keys"), keys.numel(), res.data_ptr(), _stream_ptr(stream)))
    r = res.cpu().numpy().view(np.uint64)
    return int(r[0]), int(r[1]), int(r[2]), int(r[3])


def set_param(name: str, value: int) -> None:
    _lib.check(_lib.load().b200sort_set_param(name.encode(), value))


def get_param(name: str) -> int:
    return _lib.load().b200sort_get_param(name.encode())


def profile_enable(on: bool) -> None:
    _lib.load().b200sort_profile_enable(int(on))


def profile_read(capacity: int = 4096) -> list[tuple[int, float]]:
    """Drain the per-kernel timings recorded since the last read: [(tag, ms)], tag 0 = histogram
    kernel, tag p+1 = digit pass p."""
    buf = (C.c_float * capacity)()
    tags = (C.c_int * capacity)()
    k = _lib.load().b200sort_profile_read(buf, tags, capacity)
    if k < 0:
        _lib.check(k)
    return [(int(tags[i]), float(buf[i])) for i in range(k)]


def launch_count() -> int:
    return int(_lib.load().b200sort_launch_count())


def warmup(max_n: int, pairs: bool = False) -> None:
    """Pays the one-time costs of the host-pointer entry points (buffers, staging, self test, kernel loads) now:
    b200sort_warmup."""
    _lib.check(_lib.load().b200sort_warmup(int(max_n), int(bool(pairs))))


def mgpu_shutdown() -> None:
    """Frees the per-shard device buffers of sort_by_devices / sort_pairs_by_devices (b200sort_mgpu_shutdown)."""
    _lib.check(_lib.load().b200sort_mgpu_shutdown())


def shutdown() -> None:
    _lib.load().b200sort_shutdown()
