"""Device-buffer plumbing: torch tensors carry HBM allocations and streams, nothing more."""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib


def torch():
    import torch as _torch
    return _torch


_cuda_ok = False


def require_cuda():
    """torch, after checking once that a CUDA device exists (torch.cuda.is_available() costs ~3 us
    per call and the single-query entry points used to make nine of them)."""
    global _cuda_ok
    t = torch()
    if not _cuda_ok:
        if not t.cuda.is_available():
            raise RuntimeError("npsims_b200: no CUDA device available (there is no CPU fallback)")
        _cuda_ok = True
    return t


def stream_ptr() -> int:
    return torch().cuda.current_stream().cuda_stream


class DeviceArray(np.ndarray):
    """An ndarray that remembers the HBM copy of its contents.

    `lsh.index` / `lsh.create_projections` return these so that the reference's calling
    pattern (index once with NumPy arrays, query many times) does not re-upload the table on
    every call.  Views and copies do not inherit the device buffer; mutating the array in
    place after it was uploaded is not supported (call `forget_device(arr)` first).
    """

    _nps_dev = None
    _nps_dev32 = None
    _nps_searcher = None

    def __array_finalize__(self, obj):
        self._nps_dev = None
        self._nps_dev32 = None
        self._nps_searcher = None

    def __reduce__(self):  # pickle / np.save as a plain array
        return np.asarray(self).view(np.ndarray).__reduce__()


def attach(arr: np.ndarray, dev) -> "DeviceArray":
    """`arr` as a DeviceArray that remembers its HBM copy `dev`.  While a device copy is attached the
    array is read-only: an in-place write would silently leave a stale copy on the GPU (call
    `forget_device(arr)` to drop the copy and write again)."""
    out = np.asarray(arr).view(DeviceArray)
    out._nps_dev = dev
    if dev is not None:
        _freeze(out)
    return out


def _freeze(arr) -> None:
    try:
        arr.flags.writeable = False
    except ValueError:
        pass


def forget_device(arr) -> None:
    if isinstance(arr, DeviceArray):
        arr._nps_dev = None
        arr._nps_dev32 = None
        arr._nps_searcher = None
        try:
            arr.flags.writeable = True
        except ValueError:      # a view of a read-only base
            pass


class Searcher:
    """Handle of the C-level single-query searcher (csrc/searcher.cu: per-thread slots, one CUDA graph per
    call) over a table resident in HBM.  Keeps the table tensor alive."""

    def __init__(self, table):
        t = torch()
        handle = ctypes.c_void_p()
        with t.cuda.device(table.device):
            _lib.check(_lib.raw("nps_searcher_create")(table.data_ptr(), table.shape[0], table.shape[1],
                                                       table.device.index or 0, ctypes.byref(handle)))
        self.handle = handle.value
        self.table = table
        self.words = int(table.shape[1])
        self.proj = None          # (id(projections object), tensor, ptr, B, D) of the last lsh.query
        self._destroy = _lib.raw("nps_searcher_destroy")

    def stats(self):
        calls, fallbacks, slots = ctypes.c_ulonglong(), ctypes.c_ulonglong(), ctypes.c_uint32()
        _lib.check(_lib.raw("nps_searcher_stats")(self.handle, ctypes.byref(calls), ctypes.byref(fallbacks), ctypes.byref(slots)))
        return {"calls": calls.value, "fallbacks": fallbacks.value, "slots": slots.value}

    def __del__(self):
        if getattr(self, "handle", None):
            self._destroy(self.handle)
            self.handle = None


_searcher_lock = threading.Lock()
SEARCHER_MAX_WORDS = 63


def searcher_for(hashes):
    """The single-query searcher cached on a DeviceArray table (created, and the table uploaded, on first
    use), or None when the table is not a DeviceArray / is wider than the fast path supports."""
    if not isinstance(hashes, DeviceArray):
        return None
    s = hashes._nps_searcher
    if s is not None:
        return s
    if hashes.ndim != 2 or hashes.dtype != np.uint64 or not 1 <= hashes.shape[1] <= SEARCHER_MAX_WORDS:
        return None
    with _searcher_lock:
        if hashes._nps_searcher is None:
            dev, _ = u64_table(hashes, "hamming_top_n", ndim=2)
            if dev.data_ptr() % 16:
                return None
            hashes._nps_searcher = Searcher(dev)
    return hashes._nps_searcher


def is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def u64_table(x, name: str, ndim: int | None = None):
    """(int64 CUDA tensor sharing the uint64 bits, came_from_host: bool)."""
    t = require_cuda()
    if is_tensor(x):
        if x.dtype not in (t.int64, t.uint64):
            raise TypeError(f"{name}: expected uint64, got {x.dtype}")
        if not x.is_cuda:
            x = x.cuda()
        x = x.contiguous()
        if x.dtype == t.uint64:
            x = x.view(t.int64)
        if ndim is not None and x.dim() != ndim:
            raise ValueError(f"{name}: expected {ndim} dimensions, got {x.dim()}")
        return x, False
    cached = getattr(x, "_nps_dev", None) if isinstance(x, DeviceArray) else None
    arr = np.asarray(x)
    if arr.dtype != np.uint64:
        # NumPy's gufunc dispatch for the reference's 'LL->L' loops (ufunc.c:564-565)
        raise TypeError(f"ufunc '{name}' not supported for the input types ({arr.dtype}); expected uint64")
    if ndim is not None and arr.ndim != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got {arr.ndim}")
    if cached is not None:
        if cached.device.index == t.cuda.current_device():
            return cached, True
        return cached.to(t.device("cuda", t.cuda.current_device())), True   # resident copy lives on another GPU: not cached
    arr = np.ascontiguousarray(arr)
    dev = t.from_numpy(arr.view(np.int64)).cuda()
    if isinstance(x, DeviceArray):
        x._nps_dev = dev
        _freeze(x)
    return dev, True


def to_host_u64(tensor) -> np.ndarray:
    return tensor.cpu().numpy().view(np.uint64)


def empty(shape, dtype, like=None):
    t = torch()
    return t.empty(shape, dtype=dtype, device=like.device if like is not None else "cuda")


def ptr(tensor) -> int:
    return tensor.data_ptr() if tensor is not None else 0
