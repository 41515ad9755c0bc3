"""ctypes binding of libnpsims_b200.so (C ABI: include/npsims_b200.h).

There is no CPU fallback: if the library is missing this module raises at import, and if no
CUDA device is present every compute call raises RuntimeError with the library's message.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libnpsims_b200.so")

NPS_OK, NPS_EINVAL, NPS_ECUDA, NPS_EWORKSPACE = 0, -1, -2, -3
NPS_MAX_K, NPS_MAX_WORDS = 2048, 1024
ORDER_CANONICAL, ORDER_REFERENCE = 0, 1
HASH_AUTO, HASH_TF32, HASH_BF16 = 0, 1, 2
SCAN_AUTO, SCAN_POPC, SCAN_MMA = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is not built. Run `python np-sims_b200/build.py` (needs nvcc); "
        "np_sims on B200 has no CPU fallback."
    )

_lib = ctypes.CDLL(LIB_PATH)

_vp, _u64, _u32, _int, _sz = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_size_t

_SIGS = {
    "nps_version": (_int, []),
    "nps_last_error": (ctypes.c_char_p, []),
    "nps_device_count": (_int, []),
    "nps_profile_enable": (_int, [_int]),
    "nps_profile_last": (_int, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
    "nps_last_scan_plan": (_int, [_vp]),
    "nps_launch_count": (ctypes.c_ulonglong, []),
    "nps_set_scan_path": (_int, [_int]),
    "nps_last_scan_path": (_int, []),
    "nps_set_mma_pair": (_int, [_int]),
    "nps_last_mma_plan": (_int, [_vp]),
    "nps_profile_last_phases": (_int, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_u32), _u32, ctypes.POINTER(_u32)]),
    "nps_top_n_overflow_count": (_int, [_vp, _vp, ctypes.POINTER(ctypes.c_uint32)]),
    "nps_host_hamming_top_k": (_int, [_vp, _vp, _u64, _u64, _u32, _vp]),
    "nps_host_hamming": (_int, [_vp, _vp, _u64, _u64, _vp]),
    "nps_index_create": (_int, [_vp, _u64, _u32, _int, ctypes.POINTER(_vp)]),
    "nps_index_destroy": (None, [_vp]),
    "nps_index_query_top_n": (_int, [_vp, _vp, _u32, _u32, _int, _vp, _vp]),
    "nps_index_hamming": (_int, [_vp, _vp, _vp]),
    "nps_searcher_create": (_int, [_vp, _u64, _u32, _int, ctypes.POINTER(_vp)]),
    "nps_searcher_destroy": (None, [_vp]),
    "nps_searcher_top_k": (_int, [_vp, _vp, _u32, _vp]),
    "nps_searcher_query_vector": (_int, [_vp, _vp, _int, _u32, _vp, _u32, _u32, _vp]),
    "nps_searcher_stats": (_int, [_vp, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(_u32)]),
    "nps_hamming": (_int, [_vp, _u64, _u32, _vp, _u32, _vp, _int, _vp]),
    "nps_hamming_top_n_workspace_bytes": (_sz, [_u64, _u32, _u32, _u32]),
    "nps_hamming_top_n": (_int, [_vp, _u64, _u32, _vp, _u32, _u32, _u64, _vp, _vp, _vp, _sz, _vp]),
    "nps_hamming_top_n_keys": (_int, [_vp, _u64, _u32, _vp, _u32, _u32, _u64, _vp, _vp, _sz, _vp]),
    "nps_top_n_merge_workspace_bytes": (_sz, [_u32, _u32, _u32]),
    "nps_top_n_merge": (_int, [_vp, _u32, _u32, _u32, _int, _vp, _vp, _vp, _sz, _vp]),
    "nps_hamming_top_n_reference_workspace_bytes": (_sz, [_u64, _u32, _u32, _u32]),
    "nps_hamming_top_n_reference": (_int, [_vp, _u64, _u32, _vp, _u32, _u32, _vp, _vp, _vp, _sz, _vp]),
    "nps_num_unshared_bits": (_int, [_vp, _u64, _vp, _u64, _u64, _vp, _vp]),
    "nps_rerank_top_n": (_int, [_vp, _u64, _u32, _vp, _u32, _vp, _u32, _u32, _vp, _vp, _vp]),
    "nps_project_sign_pack_workspace_bytes": (_sz, [_u64, _u32]),
    "nps_project_split_projections": (_int, [_vp, _u32, _u32, _vp, _vp]),
    "nps_project_split_bytes": (_sz, [_u32, _u32]),
    "nps_set_hash_path": (_int, [_int]),
    "nps_project_sign_pack": (_int, [_vp, _u64, _u32, _vp, _vp, _u32, ctypes.c_float, _vp, _vp, _sz, _vp]),
    "nps_project_sign_pack_stats": (_int, [_vp, _u64, _u32, _vp, ctypes.POINTER(_u32), ctypes.POINTER(_u32)]),
    "nps_project_sign_pack_stats_in_kernel": (_int, [_vp, _u64, _vp, ctypes.POINTER(_u32)]),
    "nps_project_sign_pack_f64": (_int, [_vp, _int, _u64, _u32, _vp, _u32, _vp, _vp, _vp]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(_lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTS = tuple(_SIGS)


class ScanPlan(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "tile_rows", "tile_bytes", "stages", "q_tile", "num_qtiles", "chunk_rows", "num_chunks", "cap",
        "smem_bytes", "grid", "specialised", "rows_per_thread")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class MmaPlan(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "n_tile", "num_qtiles", "cap", "growth", "a_stages", "smem_bytes", "num_phases", "last_tiles",
        "last_tiles_per_chunk", "last_num_chunks", "last_grid", "pair")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def last_mma_plan() -> dict:
    plan = MmaPlan()
    check(_lib.nps_last_mma_plan(ctypes.byref(plan)))
    return plan.as_dict()


_scan_path = SCAN_AUTO


def profile_last_phases():
    """[(tiles, scan_ms)] per phase of the last profiled tensor-core call."""
    ms = (ctypes.c_float * 32)()
    tiles = (_u32 * 32)()
    n = _u32(0)
    check(_lib.nps_profile_last_phases(ms, tiles, 32, ctypes.byref(n)))
    return [(int(tiles[i]), float(ms[i])) for i in range(n.value)]


def set_scan_path(path: int) -> int:
    """Select the scan kernel family (SCAN_AUTO / SCAN_POPC / SCAN_MMA); returns the previous one."""
    global _scan_path
    check(_lib.nps_set_scan_path(path))
    prev, _scan_path = _scan_path, path
    return prev


def overflow_count(ws_ptr: int, stream: int) -> int:
    n = ctypes.c_uint32(0)
    check(_lib.nps_top_n_overflow_count(ws_ptr, stream, ctypes.byref(n)))
    return int(n.value)


def last_scan_plan() -> dict:
    plan = ScanPlan()
    check(_lib.nps_last_scan_plan(ctypes.byref(plan)))
    return plan.as_dict()


def profile_last():
    a, b = ctypes.c_float(), ctypes.c_float()
    check(_lib.nps_profile_last(ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def last_error() -> str:
    return _lib.nps_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a C status to the exception the reference's Python surface would raise."""
    if rc == NPS_OK:
        return
    msg = last_error()
    if rc == NPS_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(f"npsims_b200: {msg} (status {rc})")


def call(name: str, *args):
    check(getattr(_lib, name)(*args))


def raw(name: str):
    return getattr(_lib, name)
