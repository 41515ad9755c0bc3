"""Sharded search across the GPUs of one box (one process per GPU, torch.distributed).

Two ways to split the work, chosen by `choose_sharding` from the table size:

* rows (ShardedIndex)      — the table is split by row, queries replicated; one all-gather of
                             candidate keys + K4 merge.  For tables that are large per GPU (C4).
* queries (ReplicatedIndex) — the table is replicated (it is small: C2 = 32 MB), each rank answers
                             its own slice of the query batch; one all-gather of finished results,
                             no merge.  Per-rank work is 1/G of the batch with no replicated part.

Row sharding in detail:

Not in the reference (single process).  The signature table shards naturally by row: rank r
holds rows [bounds[r], bounds[r+1]) and runs the same fused scan+select (K2/K3) on its shard,
emitting per-query candidate KEYS  distance << 40 | global_row  (global = local + row_base).
One NCCL all-gather moves the [q, k] keys of every rank to every rank (k*8 bytes per query
per rank — latency-, not bandwidth-bound over NVSwitch), and K4 merges the G lists.  Because
keys are totally ordered by (distance, global row), the G-GPU result is bit-identical to the
single-GPU result on the concatenated table.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._device import ptr, require_cuda, stream_ptr, u64_table


def shard_bounds(num_rows: int, world_size: int):
    """Contiguous, even row ranges: rank r owns [b[r], b[r+1]).  Boundaries are kept even so
    that every shard of an odd-width table stays 16-byte aligned inside the parent buffer."""
    b = [(num_rows * r // world_size) & ~1 for r in range(world_size)] + [num_rows]
    return b


def _cuda_local_keys(H, Q, k, row_base):
    t = require_cuda()
    m, w, q = H.shape[0], H.shape[1], Q.shape[0]
    keys = t.empty((q, k), dtype=t.int64, device=H.device)
    ws_bytes = _lib.raw("nps_hamming_top_n_workspace_bytes")(m, w, q, k)
    if ws_bytes == 0:
        raise ValueError(_lib.last_error())
    ws = t.empty(ws_bytes, dtype=t.uint8, device=H.device)
    with t.cuda.device(H.device):
        _lib.call("nps_hamming_top_n_keys", ptr(H), m, w, ptr(Q), q, k, int(row_base), ptr(keys), ptr(ws), ws_bytes,
                  stream_ptr())
        from .ufuncs import _remember_workspace
        _remember_workspace(ws)
    return keys


def _cuda_merge(gathered, k):
    """gathered [G, q, k] int64 keys (list-major, as all_gather leaves them) -> rows, dists."""
    t = require_cuda()
    G, q = gathered.shape[0], gathered.shape[1]
    rows = t.empty((q, k), dtype=t.int64, device=gathered.device)
    dists = t.empty((q, k), dtype=t.int64, device=gathered.device)
    ws_bytes = _lib.raw("nps_top_n_merge_workspace_bytes")(q, G, k)
    ws = t.empty(max(ws_bytes, 16), dtype=t.uint8, device=gathered.device)
    with t.cuda.device(gathered.device):
        _lib.call("nps_top_n_merge", ptr(gathered), G, q, k, 1, ptr(rows), ptr(dists), ptr(ws), ws_bytes,
                  stream_ptr())
    return rows, dists


class ShardedIndex:
    """This rank's shard of a row-sharded signature table.

    local_hashes : uint64 [rows_local, words] (NumPy or CUDA tensor) — this rank's rows
    row_base     : global id of local row 0
    group        : torch.distributed process group (None = default; world_size 1 if
                   torch.distributed is not initialised)
    local_keys_fn / merge_fn : injection points so the CPU (gloo) tests can drive the
                   distributed plumbing without a GPU; the defaults are the CUDA kernels.
    """

    def __init__(self, local_hashes, row_base: int, group=None, local_keys_fn=None, merge_fn=None):
        self._keys_fn = local_keys_fn or _cuda_local_keys
        self._merge_fn = merge_fn or _cuda_merge
        self._cuda = local_keys_fn is None
        self.hashes = u64_table(local_hashes, "ShardedIndex", ndim=2)[0] if self._cuda else local_hashes
        self.row_base = int(row_base)
        self.group = group

    @classmethod
    def from_global(cls, hashes, rank: int, world_size: int, **kw):
        """Slice a table every rank can see (tests, benchmarks) into this rank's shard."""
        b = shard_bounds(len(hashes), world_size)
        return cls(hashes[b[rank]:b[rank + 1]], b[rank], **kw)

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_world_size(self.group)
        return None, 1

    def query(self, query_sigs, n: int = 10):
        """query_sigs uint64 [q, words] (replicated on every rank) -> (rows, dists) [q, n] int64
        tensors carrying uint64 bits, identical on every rank."""
        import torch
        dist, world = self._world()
        Q = u64_table(query_sigs, "ShardedIndex.query", ndim=2)[0] if self._cuda else query_sigs
        keys = self._keys_fn(self.hashes, Q, int(n), self.row_base)          # [q, n]
        if world == 1:
            gathered = keys.reshape(1, *keys.shape)
        else:
            flat = torch.empty((world * keys.shape[0], keys.shape[1]), dtype=keys.dtype, device=keys.device)
            dist.all_gather_into_tensor(flat, keys.contiguous(), group=self.group)
            gathered = flat.view(world, *keys.shape)                          # list-major [G, q, k]
        return self._merge_fn(gathered, int(n))


def query_bounds(num_queries: int, world_size: int):
    """Contiguous query slices: rank r answers queries [b[r], b[r+1])."""
    return [num_queries * r // world_size for r in range(world_size)] + [num_queries]


REPLICATE_MAX_BYTES = 4 << 30


def choose_sharding(num_rows: int, words: int, num_queries: int, world_size: int) -> str:
    """"queries" when the whole table is small enough to replicate on every GPU (<= 4 GiB of the
    180 GB) and every rank still gets a tensor-core-sized slice of the batch (>= 64 queries);
    "rows" otherwise.  With a small table the row-sharded step is dominated by what does not
    shrink with G (hashing the replicated query batch, per-phase launches, the merge)."""
    if world_size <= 1:
        return "rows"
    if num_rows * words * 8 <= REPLICATE_MAX_BYTES and num_queries // world_size >= 64:
        return "queries"
    return "rows"


def _cuda_top_n(H, Q, k, return_distances=True):
    from .ufuncs import hamming_top_n
    return hamming_top_n(H, Q, n=k, return_distances=return_distances)


class ReplicatedIndex:
    """The whole table on every rank; the query batch is split across ranks.

    hashes   : uint64 [rows, words], identical on every rank
    top_n_fn : injection point for the CPU (gloo) tests; default = the CUDA top-n.
    """

    def __init__(self, hashes, group=None, top_n_fn=None):
        self._top_n = top_n_fn or _cuda_top_n
        self._cuda = top_n_fn is None
        self.hashes = u64_table(hashes, "ReplicatedIndex", ndim=2)[0] if self._cuda else hashes
        self.group = group

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_world_size(self.group), dist.get_rank(self.group)
        return None, 1, 0

    def my_slice(self, num_queries: int):
        _, world, rank = self._world()
        b = query_bounds(num_queries, world)
        return b[rank], b[rank + 1]

    def query_local(self, local_query_sigs, n: int = 10, return_distances: bool = True):
        """This rank's slice of the batch -> (rows, dists) [q_local, n] (or rows alone); no
        communication."""
        Q = u64_table(local_query_sigs, "ReplicatedIndex.query", ndim=2)[0] if self._cuda else local_query_sigs
        if return_distances:
            return self._top_n(self.hashes, Q, int(n))
        return self._top_n(self.hashes, Q, int(n), False)

    def gather(self, local, num_queries: int):
        """All-gather per-rank result slices [q_local, n] into [num_queries, n] on every rank
        (slices are padded to the longest one, ceil(q / G), for the collective)."""
        import torch
        dist, world, rank = self._world()
        if world == 1:
            return local
        out_dtype = local.dtype
        if out_dtype == torch.uint64:          # collectives move the bits as int64
            local = local.view(torch.int64)
        b = query_bounds(num_queries, world)
        longest = max(b[r + 1] - b[r] for r in range(world))
        send = local
        if local.shape[0] != longest:
            send = torch.zeros((longest, local.shape[1]), dtype=local.dtype, device=local.device)
            send[:local.shape[0]] = local
        flat = torch.empty((world * longest, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, send.contiguous(), group=self.group)
        if any(b[r + 1] - b[r] != longest for r in range(world)):
            flat = torch.cat([flat[r * longest: r * longest + b[r + 1] - b[r]] for r in range(world)])
        return flat.view(out_dtype)

    def query(self, query_sigs, n: int = 10):
        """query_sigs uint64 [q, words] replicated on every rank -> (rows, dists) [q, n], identical
        on every rank and equal to the single-GPU result (each query is answered whole)."""
        lo, hi = self.my_slice(len(query_sigs))
        rows, dists = self.query_local(query_sigs[lo:hi], n)
        return self.gather(rows, len(query_sigs)), self.gather(dists, len(query_sigs))


def keys_to_numpy(keys) -> np.ndarray:
    return keys.cpu().numpy().view(np.uint64)
