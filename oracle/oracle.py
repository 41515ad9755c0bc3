"""CPU oracle for np-sims' LSH search path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``np-sims_b200/``)
never does: its path is CUDA-only and raises if the extension is missing.

Three layers, each usable as a checker for the next:

1. ``ref_ufuncs()``  — the UNMODIFIED reference ``np_sims/ufunc.c`` compiled by
   ``oracle/Makefile`` into ``oracle/_ref/`` and loaded as the top-level module
   ``ufuncs`` (its init symbol is ``PyInit_ufuncs``, ufunc.c:579).  Exists wherever
   ``make -C oracle ref`` ran (this container; travels to the GPU box as a built file).
2. ``liboracle.so``  — ``oracle/oracle.c``, our plain-C restatement (ctypes).
3. NumPy restatements below (``np.bitwise_count``; stable argsort; float64 GEMM).

Parity status: PINNED — see tests/test_oracle.py (reference KATs + golden vectors
generated from layer 1 by tests/golden/make_golden.py).

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
UINT64_MAX = np.iinfo(np.uint64).max
_EXT = sysconfig.get_config_var("EXT_SUFFIX")


# --------------------------------------------------------------------------- build / load
def build(ref: bool = True) -> None:
    """Compile oracle.c (always) and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    if ref and os.path.exists("/root/reference/np_sims/ufunc.c"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = ctypes.CDLL(path)
        u64p, u64 = ctypes.c_void_p, ctypes.c_uint64
        L.orc_popcount64.restype = u64
        L.orc_popcount64.argtypes = [u64]
        for name in ("orc_hamming", "orc_hamming_swar"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [u64p, u64p, u64, u64, u64p]
        L.orc_num_unshared_bits.restype = None
        L.orc_num_unshared_bits.argtypes = [u64p, u64p, u64, u64p]
        for name in ("orc_hamming_top_k_queue", "orc_hamming_top_k_canonical"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [u64p, u64p, u64, u64, u64, u64p, u64p]
        L.orc_index.restype = None
        L.orc_index.argtypes = [u64p, u64, u64, u64p, ctypes.c_int, u64, u64p]
        L.orc_project.restype = None
        L.orc_project.argtypes = [u64p, u64, u64, u64p, ctypes.c_int, u64, u64p]
        L.orc_two_phase.restype = None
        L.orc_two_phase.argtypes = [u64p, u64p, u64, u64, u64, u64, u64, u64p]
        _lib = L
    return _lib


def ref_path(variant: str = "") -> str:
    """Path of the compiled reference module: variant '' (-O3), 'popcnt', 'native'."""
    return os.path.join(HERE, "_ref", variant, "ufuncs" + _EXT)


_ref_mods: dict = {}


def ref_ufuncs(variant: str = ""):
    """The reference's own gufuncs (hamming, hamming_top_10, hamming_top_cand,
    num_unshared_bits), or None if oracle/_ref was not built."""
    if variant not in _ref_mods:
        path = ref_path(variant)
        if not os.path.exists(path):
            _ref_mods[variant] = None
        else:
            spec = importlib.util.spec_from_file_location("ufuncs", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ref_mods[variant] = mod
    return _ref_mods[variant]


def ref_variant_runs(variant: str) -> bool:
    """Probe a reference build in a subprocess (a -march=native build made on another
    host CPU may die with SIGILL)."""
    path = ref_path(variant)
    if not os.path.exists(path):
        return False
    code = (
        "import importlib.util,numpy as np;"
        f"s=importlib.util.spec_from_file_location('ufuncs',{path!r});"
        "m=importlib.util.module_from_spec(s);s.loader.exec_module(m);"
        "h=np.arange(64*16,dtype=np.uint64).reshape(64,16);"
        "assert 3 in m.hamming_top_10(h,h[3]).tolist();assert m.hamming(h,h[3])[3]==0"
    )
    return subprocess.run([sys.executable, "-c", code], capture_output=True).returncode == 0


def _u64c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- distances
def hamming(hashes, query) -> np.ndarray:
    """ufunc.c:74-106 via oracle.c: uint64[m] distances of one query."""
    h, q = _u64c(hashes), _u64c(query)
    assert h.ndim == 2 and q.ndim == 1 and h.shape[1] == q.shape[0]
    out = np.empty(h.shape[0], dtype=np.uint64)
    lib().orc_hamming(_p(h), _p(q), h.shape[0], h.shape[1], _p(out))
    return out


def hamming_swar(hashes, query) -> np.ndarray:
    h, q = _u64c(hashes), _u64c(query)
    out = np.empty(h.shape[0], dtype=np.uint64)
    lib().orc_hamming_swar(_p(h), _p(q), h.shape[0], h.shape[1], _p(out))
    return out


def hamming_np(hashes, query) -> np.ndarray:
    """np_sims/hamming.py:41-61 restated with np.bitwise_count; query may be [n] or [q,n]."""
    h, q = _u64c(hashes), _u64c(query)
    if q.ndim == 1:
        return np.bitwise_count(h ^ q).sum(axis=1, dtype=np.uint64)
    return np.stack([np.bitwise_count(h ^ qq).sum(axis=1, dtype=np.uint64) for qq in q])


def num_unshared_bits(a, b) -> np.ndarray:
    a, b = np.broadcast_arrays(_u64c(a), _u64c(b))
    a, b = _u64c(a), _u64c(b)
    out = np.empty(a.shape, dtype=np.uint8)
    lib().orc_num_unshared_bits(_p(a), _p(b), a.size, _p(out))
    return out


# --------------------------------------------------------------------------- top-n
def top_k_queue(hashes, query, k: int):
    """Reference queue semantics for any k (ufunc.c:108-195, :411-441).
    Returns (rows[k] in slot order, scores[k]); unfilled slots are UINT64_MAX."""
    h, q = _u64c(hashes), _u64c(query)
    if h.ndim != 2:
        h = h.reshape(0, q.shape[0])
    rows = np.empty(k, dtype=np.uint64)
    scores = np.empty(k, dtype=np.uint64)
    lib().orc_hamming_top_k_queue(_p(h), _p(q), h.shape[0], h.shape[1], k, _p(rows), _p(scores))
    return rows, scores


def top_k_canonical(hashes, query, k: int):
    """Canonical contract: k smallest (distance, row), ascending (north_star:
    'ties broken by lowest index').  Returns (rows[k], dists[k]), UINT64_MAX padded."""
    h, q = _u64c(hashes), _u64c(query)
    if h.ndim != 2:
        h = h.reshape(0, q.shape[0])
    rows = np.empty(k, dtype=np.uint64)
    dists = np.empty(k, dtype=np.uint64)
    lib().orc_hamming_top_k_canonical(_p(h), _p(q), h.shape[0], h.shape[1], k, _p(rows), _p(dists))
    return rows, dists


def top_k_canonical_np(hashes, query, k: int):
    """Same contract via NumPy: stable argsort of the distances (np_sims/lsh.py:54 with a
    stable sort)."""
    d = hamming_np(hashes, query)
    order = np.argsort(d, kind="stable")[:k]
    rows = np.full(k, UINT64_MAX, dtype=np.uint64)
    dists = np.full(k, UINT64_MAX, dtype=np.uint64)
    rows[: len(order)] = order
    dists[: len(order)] = d[order]
    return rows, dists


def tie_aware_equal(dist_all: np.ndarray, ids_a, ids_b, k: int) -> bool:
    """SURVEY.md §8c comparator between two top-k id sets that may break boundary ties
    differently: same multiset of distances, every id within the k-th distance, and every
    row strictly better than the k-th distance present in both."""
    a = np.asarray([i for i in np.asarray(ids_a, dtype=np.uint64) if i != UINT64_MAX], dtype=np.int64)
    b = np.asarray([i for i in np.asarray(ids_b, dtype=np.uint64) if i != UINT64_MAX], dtype=np.int64)
    n = len(dist_all)
    if len(a) != min(k, n) or len(b) != min(k, n):
        return False
    if len(set(a.tolist())) != len(a) or len(set(b.tolist())) != len(b):
        return False
    if len(a) == 0:
        return True
    da, db = np.sort(dist_all[a]), np.sort(dist_all[b])
    if not np.array_equal(da, db):
        return False
    kth = np.sort(dist_all, kind="stable")[len(a) - 1]
    if da[-1] > kth:
        return False
    strictly_better = set(np.nonzero(dist_all < kth)[0].tolist())
    return strictly_better <= set(a.tolist()) and strictly_better <= set(b.tolist())


# --------------------------------------------------------------------------- hashing
def index(vectors, projections) -> np.ndarray:
    """np_sims/lsh.py:36-44 via oracle.c: uint64[N, B/64] signatures, float64 arithmetic."""
    projections = np.ascontiguousarray(projections, dtype=np.float64)
    if len(projections) % 64 != 0:
        raise ValueError("Projections must be a multiple of 64")
    vectors = np.asarray(vectors)
    is_f32 = vectors.dtype == np.float32
    vectors = np.ascontiguousarray(vectors, dtype=np.float32 if is_f32 else np.float64)
    if vectors.ndim == 1:
        vectors = vectors[None, :]
    out = np.empty((vectors.shape[0], len(projections) // 64), dtype=np.uint64)
    lib().orc_index(_p(projections), projections.shape[0], projections.shape[1],
                    _p(vectors), int(is_f32), vectors.shape[0], _p(out))
    return out


def index_np(vectors, projections) -> np.ndarray:
    """np_sims/lsh.py:31-44 restated as one float64 GEMM + packbits (bit-identical to the
    per-row loop on the probe data)."""
    projections = np.asarray(projections, dtype=np.float64)
    if len(projections) % 64 != 0:
        raise ValueError("Projections must be a multiple of 64")
    dots = np.asarray(vectors, dtype=np.float64) @ projections.T
    bits = np.packbits(dots >= 0, axis=-1, bitorder="little")
    return np.ascontiguousarray(bits).view(np.uint64)


def project(vectors, projections) -> np.ndarray:
    """float64 dot products x.w (for the |x.w| < tol hash tolerance check)."""
    projections = np.ascontiguousarray(projections, dtype=np.float64)
    vectors = np.asarray(vectors)
    is_f32 = vectors.dtype == np.float32
    vectors = np.ascontiguousarray(vectors, dtype=np.float32 if is_f32 else np.float64)
    out = np.empty((vectors.shape[0], projections.shape[0]), dtype=np.float64)
    lib().orc_project(_p(projections), projections.shape[0], projections.shape[1],
                      _p(vectors), int(is_f32), vectors.shape[0], _p(out))
    return out


def create_projections(num_projections: int, dims: int) -> np.ndarray:
    """np_sims/lsh.py:5-28: Gaussian rows from the GLOBAL np.random state, unit-normalised."""
    rows = []
    for _ in range(num_projections):
        p = np.random.normal(size=dims)
        p /= np.linalg.norm(p)
        rows.append(p)
    return np.array(rows)


def two_phase(hashes, query, front_words: int, cand: int = 1000, k: int = 10) -> np.ndarray:
    """np_sims/lsh.py:84-91 on signatures."""
    h, q = _u64c(hashes), _u64c(query)
    out = np.empty(k, dtype=np.uint64)
    lib().orc_two_phase(_p(h), _p(q), h.shape[0], h.shape[1], front_words, cand, k, _p(out))
    return out


def two_phase_canonical(hashes, query, front_words: int, cand: int = 1000, k: int = 10):
    """np_sims/lsh.py:84-91 with the canonical tie rule at BOTH selections (the batched
    `lsh.query_two_phase_batch`): candidates = the `cand` rows with the smallest distance on the
    first `front_words` words, ties by lowest row id; result = the k candidates with the smallest
    full-width distance, ties by lowest row id.  -> (rows, dists), UINT64_MAX padded."""
    h, q = _u64c(hashes), _u64c(query)
    front = hamming_np(np.ascontiguousarray(h[:, :front_words]), q[:front_words])
    c = np.argsort(front, kind="stable")[:cand]
    full = hamming_np(h[c], q)
    order = np.lexsort((c, full))[:k]
    rows = np.full(k, np.iinfo(np.uint64).max, dtype=np.uint64)
    dists = np.full(k, np.iinfo(np.uint64).max, dtype=np.uint64)
    rows[:len(order)] = c[order].astype(np.uint64)
    dists[:len(order)] = full[order].astype(np.uint64)
    return rows, dists
