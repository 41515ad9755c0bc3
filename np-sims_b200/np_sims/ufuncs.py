"""B200 replacements for the gufuncs of the reference's native module `np_sims.ufuncs`
(/root/reference/np_sims/ufunc.c:591-627): `hamming`, `hamming_top_10`, `hamming_top_cand`,
`num_unshared_bits` — same names, same positional `(hashes, query)` signature — plus the
general batched `hamming_top_n` the north-star asks for.

Inputs may be NumPy arrays (uploaded; results come back as NumPy) or torch CUDA tensors of
dtype int64/uint64 carrying the uint64 bits (results stay on the device).  `query` may be
`[n]` (reference-compatible, result `[k]`) or `[q, n]` (batched, result `[q, k]`).

Every function launches the hand-written sm_100a kernels through the C ABI; nothing here
computes on the CPU.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._device import DeviceArray, empty, is_tensor, ptr, require_cuda, searcher_for, stream_ptr, to_host_u64, u64_table

UINT64_MAX = np.iinfo(np.uint64).max
_searcher_top_k = _lib.raw("nps_searcher_top_k")


def _prepare(hashes, query, name):
    """-> (H dev [m,n], Q dev [q,n], single_query, host_result)"""
    H, h_host = u64_table(hashes, name, ndim=2)
    Q, q_host = u64_table(query, name)
    if Q.dim() not in (1, 2):
        raise ValueError(f"{name}: query must be [n] or [q, n], got {tuple(Q.shape)}")
    single = Q.dim() == 1
    if single:
        Q = Q.reshape(1, -1)
    if H.shape[1] != Q.shape[1]:
        # the message NumPy gives for the reference's "(m,n),(n)->..." signatures
        raise ValueError(
            f"{name}: Input operand 1 has a mismatch in its core dimension 0, with gufunc "
            f"signature (m,n),(n)->(k) (size {Q.shape[1]} is different from {H.shape[1]})"
        )
    if H.shape[1] == 0:
        raise ValueError(f"{name}: signatures must have at least one 64-bit word")
    if Q.device != H.device:
        Q = Q.to(H.device)
    host = not (is_tensor(hashes) or is_tensor(query))
    return H, Q, single, host


def _finish(t_out, single, host, as_u64=True):
    if single:
        t_out = t_out[0]
    if host:
        return to_host_u64(t_out) if as_u64 else t_out.cpu().numpy()
    if as_u64:
        return t_out.view(require_cuda().uint64)
    return t_out


def hamming(hashes, query, out_dtype=np.uint64):
    """gufunc `hamming`, "(m,n),(n)->(m)" (ufunc.c:74-106, registered :599-602; exported as
    `hamming_c`, np_sims/__init__.py:2): out[i] = sum_w popcount(hashes[i,w] ^ query[w]).
    Batched `[q,n]` queries give `[q,m]`.  `out_dtype` uint64 (reference), uint32 or uint16."""
    t = require_cuda()
    H, Q, single, host = _prepare(hashes, query, "hamming")
    nbytes = np.dtype(out_dtype).itemsize
    tdtype = {8: t.int64, 4: t.int32, 2: t.int16}[nbytes]
    out = t.empty((Q.shape[0], H.shape[0]), dtype=tdtype, device=H.device)
    with t.cuda.device(H.device):
        _lib.call("nps_hamming", ptr(H), H.shape[0], H.shape[1], ptr(Q), Q.shape[0], ptr(out), nbytes, stream_ptr())
    if single:
        out = out[0]
    if host:
        return out.cpu().numpy().view(np.dtype(out_dtype))
    return out.view({8: t.uint64, 4: t.uint32, 2: t.uint16}[nbytes])


def hamming_top_n(hashes, query, n=10, return_distances=False, order="canonical", row_base=0):
    """Top-n rows by Hamming distance.

    order="canonical": ascending (distance, row id) — "ties broken by lowest index"; equals
        np.argsort(hamming(hashes, query), kind="stable")[:n].  Fused scan+select kernel (K2)
        followed by the block merge (K3).
    order="reference": bit-identical to the reference queue (ufunc.c:150-195): ids in SLOT
        order, tie survivors as the sequential insert/evict history leaves them.

    Rows beyond the table (fewer than n rows) are padded with UINT64_MAX, like ufunc.c:129-132.
    Returns ids `[n]`/`[q,n]` uint64, or `(ids, distances)` if return_distances."""
    t = require_cuda()
    if order not in ("canonical", "reference"):
        raise ValueError("order must be 'canonical' or 'reference'")
    n = int(n)
    if not 1 <= n <= _lib.NPS_MAX_K:
        raise ValueError(f"n={n} outside [1, {_lib.NPS_MAX_K}]")
    if (order == "reference" and not return_distances and not row_base and type(query) is np.ndarray
            and query.ndim == 1 and query.dtype == np.uint64 and isinstance(hashes, DeviceArray)
            and hashes.ndim == 2 and hashes.shape[1] == query.shape[0]):
        # the reference's calling pattern (one NumPy query per call, often from many threads) over a
        # resident table: one C call = one CUDA-graph launch on the calling thread's own stream
        s = searcher_for(hashes)
        if s is not None:
            q = query if query.flags.c_contiguous else np.ascontiguousarray(query)
            out = np.empty(n, dtype=np.uint64)
            _lib.check(_searcher_top_k(s.handle, q.ctypes.data, n, out.ctypes.data))
            return out
    H, Q, single, host = _prepare(hashes, query, "hamming_top_n")
    m, w, q = H.shape[0], H.shape[1], Q.shape[0]
    rows = t.empty((q, n), dtype=t.int64, device=H.device)
    dists = t.empty((q, n), dtype=t.int64, device=H.device) if return_distances else None
    with t.cuda.device(H.device):
        if order == "canonical":
            ws_bytes = _lib.raw("nps_hamming_top_n_workspace_bytes")(m, w, q, n)
            if ws_bytes == 0:
                raise ValueError(_lib.last_error())
            ws = t.empty(ws_bytes, dtype=t.uint8, device=H.device)
            _lib.call("nps_hamming_top_n", ptr(H), m, w, ptr(Q), q, n, int(row_base), ptr(rows), ptr(dists),
                      ptr(ws), ws_bytes, stream_ptr())
            _remember_workspace(ws)
        else:
            if row_base:
                raise ValueError("row_base is only supported with order='canonical'")
            # [batch, m] distances live in the workspace: bound the batch to ~1 GiB
            per_q = max(1, _lib.raw("nps_hamming_top_n_reference_workspace_bytes")(m, w, 1, n))
            batch = max(1, min(q, (1 << 30) // per_q))
            ws_bytes = _lib.raw("nps_hamming_top_n_reference_workspace_bytes")(m, w, batch, n)
            ws = t.empty(max(ws_bytes, 16), dtype=t.uint8, device=H.device)
            for q0 in range(0, q, batch):
                nq = min(batch, q - q0)
                _lib.call("nps_hamming_top_n_reference", ptr(H), m, w, ptr(Q[q0:]), nq, n, ptr(rows[q0:]),
                          ptr(dists[q0:]) if dists is not None else 0, ptr(ws), ws_bytes, stream_ptr())
    if return_distances:
        return _finish(rows, single, host), _finish(dists, single, host)
    return _finish(rows, single, host)


def hamming_rerank(hashes, query, candidates, n=10, return_distances=False):
    """Second phase of the two-phase search (lsh.py:88-91): of each query's `candidates` (row ids
    `[c]` / `[q, c]`, e.g. `hamming_top_n` on the front words) the n with the smallest FULL-width
    distance, ascending (distance, row id).  The candidate rows are read in place (K6,
    rerank.cu) — no gathered copy of the table.  Ids of UINT64_MAX or beyond the table are ignored."""
    t = require_cuda()
    n = int(n)
    if not 1 <= n <= _lib.NPS_MAX_K:
        raise ValueError(f"n={n} outside [1, {_lib.NPS_MAX_K}]")
    H, Q, single, host = _prepare(hashes, query, "hamming_rerank")
    C, c_host = u64_table(candidates, "hamming_rerank")
    if C.dim() == 1:
        C = C.reshape(1, -1)
    if C.dim() != 2 or C.shape[0] != Q.shape[0] or not 1 <= C.shape[1] <= 4096:
        raise ValueError(f"hamming_rerank: candidates must be [q, c] with c <= 4096, got {tuple(C.shape)}")
    C = C.to(H.device).contiguous()
    host = host and not is_tensor(candidates)
    q = Q.shape[0]
    rows = t.empty((q, n), dtype=t.int64, device=H.device)
    dists = t.empty((q, n), dtype=t.int64, device=H.device) if return_distances else None
    with t.cuda.device(H.device):
        _lib.call("nps_rerank_top_n", ptr(H), H.shape[0], H.shape[1], ptr(Q), q, ptr(C), C.shape[1], n, ptr(rows),
                  ptr(dists), stream_ptr())
    if return_distances:
        return _finish(rows, single, host), _finish(dists, single, host)
    return _finish(rows, single, host)


_last_ws = None


_KEEP_WHOLE_BYTES = 64 << 20


def _remember_workspace(ws):
    """Keep what `last_overflow_count` needs: the overflow counter is the first word of the workspace.
    Small workspaces (single queries) are kept whole — no extra launch; of large ones (a C4 batch needs
    2 GB) only a 4-byte copy survives the call."""
    global _last_ws
    _last_ws = ws if ws.numel() <= _KEEP_WHOLE_BYTES else ws[:4].clone()


def last_overflow_count() -> int:
    """Diagnostics: how many candidate-buffer overflows the last tensor-core top-n call saw.  A
    non-zero count means the library repaired the call ON THE DEVICE by re-running it on the
    XOR+POPC scan (gated launches in api.cu: run_top_n) — results are exact either way.  Syncs."""
    if _last_ws is None or _lib.raw("nps_last_scan_path")() != _lib.SCAN_MMA:
        return 0
    return _lib.overflow_count(ptr(_last_ws), stream_ptr())


def hamming_top_10(hashes, query):
    """gufunc `hamming_top_10`, "(m,n),(n)->(10)" (ufunc.c:444-496, registered :607-610).
    Bit-identical to the reference: uint64[10] row ids in queue-slot order, UINT64_MAX padded."""
    return hamming_top_n(hashes, query, 10, order="reference")


def hamming_top_cand(hashes, query):
    """gufunc `hamming_top_cand`, "(m,n),(n)->(1000)" (ufunc.c:499-551, registered :615-618)."""
    return hamming_top_n(hashes, query, 1000, order="reference")


def num_unshared_bits(a, b):
    """ufunc `num_unshared_bits`, elementwise popcount(a ^ b) -> uint8 with broadcasting
    (ufunc.c:41-66, registered :591-593).  The reference accumulates into an uninitialised
    output (`+=`, :58); this returns the popcount itself."""
    t = require_cuda()
    host = not (is_tensor(a) or is_tensor(b))

    def dev(x, name):
        if is_tensor(x):
            return u64_table(x, name)[0]
        arr = np.asarray(x)
        if arr.dtype != np.uint64:
            raise TypeError(f"ufunc 'num_unshared_bits' not supported for the input types ({arr.dtype})")
        return u64_table(arr, name)[0]

    A, B = dev(a, "num_unshared_bits"), dev(b, "num_unshared_bits")
    if B.device != A.device:
        B = B.to(A.device)
    shape = t.broadcast_shapes(tuple(A.shape), tuple(B.shape))
    A = A.broadcast_to(shape).contiguous()
    B = B.broadcast_to(shape).contiguous()
    out = t.empty(shape, dtype=t.uint8, device=A.device)
    n = out.numel()
    with t.cuda.device(A.device):
        _lib.call("nps_num_unshared_bits", ptr(A), max(n, 1), ptr(B), max(n, 1), n, ptr(out), stream_ptr())
    return out.cpu().numpy() if host else out
