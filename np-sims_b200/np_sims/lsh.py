"""`np_sims.lsh` of the reference (/root/reference/np_sims/lsh.py) on the B200: same function
names, arguments, return conventions and error behaviour; hashing and search run as CUDA
kernels (K1 projection+sign+pack, K2/K3 scan+top-n, K5 distances, queue replay).

Line citations are to /root/reference/np_sims/lsh.py.
"""
from __future__ import annotations

import numpy as np

from np_sims import hamming_c, hamming_naive, hamming_top_10, hamming_top_cand, hamming_top_n
from . import _lib
from ._device import DeviceArray, attach, is_tensor, ptr, require_cuda, searcher_for, stream_ptr, to_host_u64


_searcher_query_vector = _lib.raw("nps_searcher_query_vector")


def random_projection(dims):
    """lsh.py:5-18 — a unit vector from the GLOBAL np.random state (Gaussian, normalised).
    Kept on the host in NumPy so that seeds reproduce the reference bit for bit."""
    projection = np.random.normal(size=dims)
    projection /= np.linalg.norm(projection)
    return projection


def create_projections(num_projections=0, dims=0) -> np.ndarray:
    """lsh.py:21-28 — float64 [num_projections, dims].  Returned as a DeviceArray so the
    first hashing call uploads it once."""
    projs = [random_projection(dims) for _ in range(num_projections)]
    return attach(np.array(projs), None)


# ------------------------------------------------------------------------------- hashing
def _projections_device(projections):
    t = require_cuda()
    if is_tensor(projections):
        p = projections
        if p.dtype != t.float64:
            p = p.to(t.float64)
        return p.cuda().contiguous()
    cached = getattr(projections, "_nps_dev", None) if isinstance(projections, DeviceArray) else None
    if cached is not None:
        return cached
    arr = np.ascontiguousarray(np.asarray(projections), dtype=np.float64)
    dev = t.from_numpy(arr).cuda()
    if isinstance(projections, DeviceArray):
        projections._nps_dev = dev
        from ._device import _freeze
        _freeze(projections)
    return dev


def _vectors_device(vectors):
    """float32 stays float32 (NumPy upcasts it inside np.dot, lsh.py:32); everything else is
    float64 like NumPy's own promotion."""
    t = require_cuda()
    if is_tensor(vectors):
        v = vectors
        if v.dtype not in (t.float32, t.float64):
            v = v.to(t.float64)
        return v.cuda().contiguous()
    arr = np.asarray(vectors)
    arr = np.ascontiguousarray(arr, dtype=np.float32 if arr.dtype == np.float32 else np.float64)
    return t.from_numpy(arr).cuda()


# Hashing statistics of the last call: path and number of (row, projection) pairs; `hash_stats()`
# adds how many pairs the tensor-core pass sent to the float64 fix-up (synchronises).
last_hash_stats = {"path": None, "pairs": 0, "recomputed": None}
_last_hash_ws = None


def hash_stats():
    """Diagnostics (tests / bench): `last_hash_stats` with `recomputed` filled in — the number of
    pairs whose accumulator was too close to zero to trust and that were recomputed in float64, by the
    fixer warps inside the hashing kernel (`in_kernel`) or through the fix-up list (`listed`);
    `overflowed` says the list was too small and the whole call was redone in float64 on
    the device.  Synchronises the stream; the hashing call itself never does."""
    st = dict(last_hash_stats)
    if st["path"] == "tc-split+fixup" and _last_hash_ws is not None:
        counter, cap = _last_hash_ws
        listed, in_kernel = (int(v) for v in counter.cpu().numpy().view(np.uint32)[:2])
        st.update(recomputed=listed + in_kernel, listed=listed, in_kernel=in_kernel, overflowed=listed > cap)
    return st


def _remember_hash_workspace(ws, ws_bytes, n):
    """Keep what `hash_stats` needs from the K1T workspace (layout, project_tc.cu: |x| norms, then the
    fix-up counters, then the list): an 8-byte copy of the counters and the list capacity — not the
    workspace itself, which is 10 GB for the C3 table."""
    global _last_hash_ws
    off = (4 * n + 255) // 256 * 256
    _last_hash_ws = (ws[off:off + 8].clone(), (ws_bytes - off - 256) // 8)


def _projections_split(projections, P64):
    """Both operand encodings of the projections (nps_project_split_projections: TF32 pair as
    float32 [2, B, D], BF16 pair as bfloat16 [2, B, Dp]) and
    max_j |w_j| for the tensor-core pass, cached on the DeviceArray next to the float64 upload."""
    t = require_cuda()
    cached = getattr(projections, "_nps_dev32", None) if isinstance(projections, DeviceArray) else None
    if cached is not None:
        return cached
    split = t.empty(_lib.raw("nps_project_split_bytes")(P64.shape[0], P64.shape[1]), dtype=t.uint8, device=P64.device)
    with t.cuda.device(P64.device):
        _lib.call("nps_project_split_projections", ptr(P64), P64.shape[0], P64.shape[1], ptr(split), stream_ptr())
    norm_max = float(t.linalg.vector_norm(P64, dim=1).max().item()) * (1.0 + 1e-6)
    out = (split, norm_max)
    if isinstance(projections, DeviceArray):
        projections._nps_dev32 = out
    return out


def hash_vectors(vectors, projections, return_dots=False, exact_only=False, tf32_raw=False):
    """Device-level hashing: vectors [N, D] -> int64 CUDA tensor [N, B/64] carrying the uint64
    signatures (bit j = (projections[j] . x >= 0) at word j//64, bit j%64 — lsh.py:33).

    float32 vectors with D % 4 == 0 take the tensor-core path (K1T: TMA-fed tcgen05 GEMM on
    two-term BF16 (or TF32, `_lib.HASH_TF32`) splits of both operands, sign/bit-pack epilogue,
    float64 fix-up of the ~0.04 % of pairs whose accumulator is too close to zero); everything else, `return_dots` and
    `exact_only` take the float64 CUDA-core kernel (K1a).  Both give the signatures of the
    reference's float64 arithmetic.  The call is asynchronous (no host sync).
    `tf32_raw=True` (measurement only) disables the fix-up, returning the bare tensor-core signs,
    so that the flipped-bit rate of the tensor-core arithmetic itself can be reported."""
    global _last_hash_ws
    t = require_cuda()
    P = _projections_device(projections)
    X = _vectors_device(vectors)
    if X.dim() == 1:
        X = X.reshape(1, -1)
    if P.dim() != 2 or X.dim() != 2 or P.shape[1] != X.shape[1]:
        raise ValueError(f"shapes {tuple(P.shape)} and {tuple(X.shape)} not aligned")
    if P.shape[0] % 64 != 0:
        raise ValueError("Projections must be a multiple of 64")
    X = X.to(P.device)
    n, d, b = X.shape[0], X.shape[1], P.shape[0]
    out = t.empty((n, b // 64), dtype=t.int64, device=P.device)
    with t.cuda.device(P.device):
        if (X.dtype == t.float32 and not return_dots and not exact_only and n > 0 and d % 4 == 0 and b <= 65536
                and X.data_ptr() % 16 == 0):
            split, norm_max = _projections_split(projections, P)
            ws_bytes = _lib.raw("nps_project_sign_pack_workspace_bytes")(n, b)
            ws = t.empty(ws_bytes, dtype=t.uint8, device=P.device)
            _lib.call("nps_project_sign_pack", ptr(X), n, d, ptr(split), ptr(P), b, 1e-30 if tf32_raw else norm_max,
                      ptr(out), ptr(ws), ws_bytes, stream_ptr())
            _remember_hash_workspace(ws, ws_bytes, n)
            last_hash_stats.update(path="tc-split+fixup", pairs=n * b, recomputed=None)
            return out
        dots = t.empty((n, b), dtype=t.float64, device=P.device) if return_dots else None
        _lib.call("nps_project_sign_pack_f64", ptr(X), int(X.dtype == t.float32), n, d, ptr(P), b, ptr(out), ptr(dots),
                  stream_ptr())
        _last_hash_ws = None
        last_hash_stats.update(path="f64", pairs=n * b, recomputed=n * b)
    return (out, dots) if return_dots else out


def index_one(query_vector, projections):
    """lsh.py:31-33 — one vector -> uint64[B/64]."""
    sig = hash_vectors(query_vector, projections)[0]
    if is_tensor(query_vector):
        return sig.view(require_cuda().uint64)
    return to_host_u64(sig)


def index(vectors, projections):
    """lsh.py:36-44 — all vectors -> uint64[N, B/64] (one batched kernel instead of the
    reference's per-row Python loop; the progress prints of :42-43 are dropped).  The NumPy
    result keeps its HBM copy attached (DeviceArray), so queries do not re-upload it."""
    if len(projections) % 64 != 0:
        raise ValueError("Projections must be a multiple of 64")
    sigs = hash_vectors(vectors, projections)
    if is_tensor(vectors):
        return sigs.view(require_cuda().uint64)
    return attach(to_host_u64(sigs), sigs)


# ------------------------------------------------------------------------------- queries
def _check_hamming_func(hamming_func):
    if hamming_func not in (hamming_c, hamming_naive):
        raise TypeError("np_sims (B200): hamming_func must be np_sims.hamming_c or np_sims.hamming_naive")


def _as_index_dtype(ids):
    # np.argsort returns intp indices (lsh.py:54); keep that dtype for drop-in use as an index
    return ids.view(np.int64) if isinstance(ids, np.ndarray) else ids


def _query_signature(vector, projections):
    """Signature of ONE query vector, left on the device: hashing feeds the scan without the
    D2H + H2D hop (and the synchronisation) that returning it as NumPy in between would cost."""
    return hash_vectors(vector, projections)[0]


def _results_like_inputs(out, vector, hashes):
    """NumPy in -> NumPy out (the reference's contract); tensors stay tensors."""
    if is_tensor(vector) or is_tensor(hashes):
        return out
    return to_host_u64(out) if is_tensor(out) else out


def query_with_hamming_then_slow_argsort(vector, hashes, projections, hamming_func=hamming_c, n=10):
    """lsh.py:47-56 — distances then argsort()[:n]; returns (idxs, distances).  Here the
    selection is the stable (distance, row) order, done on the device."""
    _check_hamming_func(hamming_func)
    query_hash = _query_signature(vector, projections)
    idxs, sims = hamming_top_n(hashes, query_hash, n=min(int(n), max(len(hashes), 1)), return_distances=True)
    keep = min(int(n), len(hashes))
    idxs, sims = _results_like_inputs(idxs, vector, hashes), _results_like_inputs(sims, vector, hashes)
    return _as_index_dtype(idxs[:keep]), sims[:keep]


def query_pure_python(vector, hashes, projections, hamming_func=hamming_naive, n=10):
    """lsh.py:59-67 — same contract as above (the reference's NumPy-only variant)."""
    return query_with_hamming_then_slow_argsort(vector, hashes, projections, hamming_func=hamming_func, n=n)


def query_with_hamming_top_n(vector, hashes, projections, hamming_func=hamming_c, n=10):
    """lsh.py:70-74 — (hamming_top_10 ids, None); `n` and `hamming_func` are ignored by the
    reference and therefore here.  A NumPy vector against a resident table (what `lsh.index` returns)
    is ONE C call: host vector in, float64 projection + sign + pack, K2R search, host ids out, all in
    one CUDA graph on the calling thread's own stream (csrc/searcher.cu)."""
    if (type(vector) is np.ndarray and vector.ndim == 1 and vector.dtype in (np.float32, np.float64)
            and isinstance(hashes, DeviceArray)):
        s = searcher_for(hashes)
        if s is not None:
            pc = s.proj
            if pc is None or pc[0] is not projections:
                P = _projections_device(projections)
                if P.dim() != 2 or P.device != s.table.device:
                    raise ValueError("projections must be a [num_projections, dims] matrix on the table's device")
                pc = s.proj = (projections, P, P.data_ptr(), int(P.shape[0]), int(P.shape[1]))
            if vector.shape[0] != pc[4]:
                raise ValueError(f"shapes {(pc[3], pc[4])} and {tuple(vector.shape)} not aligned")
            if pc[3] == 64 * s.words:
                v = vector if vector.flags.c_contiguous else np.ascontiguousarray(vector)
                out = np.empty(10, dtype=np.uint64)
                _lib.check(_searcher_query_vector(s.handle, v.ctypes.data, int(v.dtype == np.float32), pc[4], pc[2],
                                                  pc[3], 10, out.ctypes.data))
                return out, None
    query_hash = _query_signature(vector, projections)
    best = hamming_top_10(hashes, query_hash)
    return _results_like_inputs(best, vector, hashes), None


def front_hashes(hashes, front_bits=64):
    """lsh.py:77-81 — the first front_bits//64 words of every signature, as a new table."""
    front_len_64 = front_bits // 64
    if is_tensor(hashes):
        return hashes[:, 0:front_len_64].contiguous()
    dev = getattr(hashes, "_nps_dev", None) if isinstance(hashes, DeviceArray) else None
    front = np.asarray(hashes)[:, 0:front_len_64].copy()
    return attach(front, dev[:, 0:front_len_64].contiguous() if dev is not None else None)


def query_with_hamming_top_n_two_phase(vector, front_hashes, hashes, projections, hamming_func=hamming_c, n=10):
    """lsh.py:84-91 — top-1000 on the front words, gather, top-10 on the full signature, map
    back.  Both selections follow the reference queue order; UINT64_MAX padding ids index the
    last row exactly as NumPy's wrap-around does in the reference (`hashes[candidates]`)."""
    query_hash = index_one(vector, projections)
    query_hash_front = query_hash[:front_hashes.shape[1]]
    candidates = hamming_top_cand(front_hashes, query_hash_front)
    gathered = _gather_rows(hashes, candidates)
    best = hamming_top_10(gathered, query_hash)        # device tensor (gathered lives in HBM)
    t = require_cuda()
    if is_tensor(candidates):
        return candidates.view(t.int64)[best.view(t.int64)].view(t.uint64), None
    return candidates[best.view(t.int64).cpu().numpy()], None


def _gather_rows(hashes, candidates):
    """hashes[candidates] on the device (ids of UINT64_MAX wrap to the last row)."""
    t = require_cuda()
    from ._device import u64_table
    H, _ = u64_table(hashes, "hamming_top_10", ndim=2)
    C, _ = u64_table(candidates, "hamming_top_10")
    return H[C.to(H.device)]   # int64 view of UINT64_MAX is -1 -> last row


query = query_with_hamming_top_n  # Default query to fastest
query_two_phase = query_with_hamming_top_n_two_phase


def query_two_phase_batch(vectors, front_hashes, hashes, projections, n=10, num_candidates=1000,
                          return_distances=False):
    """Not in the reference as a batch: its two-phase recipe (lsh.py:84-91, the fastest-recall
    configuration of benchmark/glove.py) for a whole batch of query vectors, fused on the device —
    K1 hashes the batch, the batched scan (K2T/K2) takes the `num_candidates` best rows on the
    FRONT words, K6 re-ranks them on the full signatures in place (no gathered copy) — in
    canonical order at both stages (ascending (distance, row id)).  `query_two_phase` remains the
    single-query, reference-queue-order form."""
    sigs = hash_vectors(vectors, projections)
    t = require_cuda()
    from ._device import u64_table
    F, _ = u64_table(front_hashes, "query_two_phase_batch", ndim=2)
    front_sigs = sigs[:, :F.shape[1]].contiguous()
    cand = hamming_top_n(F, front_sigs, n=int(num_candidates))
    from np_sims.ufuncs import hamming_rerank
    H, _ = u64_table(hashes, "query_two_phase_batch", ndim=2)
    out = hamming_rerank(H, sigs, cand, n=n, return_distances=return_distances)
    if is_tensor(vectors) or is_tensor(hashes):
        return out

    def host(x):
        return x.cpu().numpy().view(np.uint64) if is_tensor(x) else x
    return tuple(host(o) for o in out) if return_distances else host(out)


# ------------------------------------------------------------------------------- batched API
def query_batch(vectors, hashes, projections, n=10, return_distances=False):
    """Not in the reference: hash a batch of query vectors (K1) and search them in one fused
    scan (K2/K3), canonical order.  This is the throughput path the benchmark measures."""
    sigs = hash_vectors(vectors, projections)
    out = hamming_top_n(hashes, sigs, n=n, return_distances=return_distances)
    if is_tensor(vectors) or is_tensor(hashes):
        return out

    def host(x):
        return x.cpu().numpy().view(np.uint64) if is_tensor(x) else x
    return tuple(host(o) for o in out) if return_distances else host(out)
