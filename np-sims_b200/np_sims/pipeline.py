"""Streaming search: host batches in, host ids out, copies overlapped with the search.

Not in the reference (its benchmark loops over single queries, benchmark/glove.py:183-199).  A
serving loop that answers one batch of query vectors after another pays, per batch, a
host-to-device copy of the vectors (12 MB for 10 000 x 300 float32), the search, and a
device-to-host copy of the ids.  `QueryStream` runs the three on three CUDA streams with `depth`
batches in flight, so that the copy of batch i + 1 and the read-back of batch i - 1 overlap the
search of batch i:

    stream = QueryStream(hashes, projections, n=10, depth=2)
    for ids in stream.map(batches):          # batches: iterable of float32 [q, dims] host arrays
        ...                                  # ids: uint64 [q, n] NumPy view of a pinned buffer,
                                             # valid until `depth` more batches were submitted

Every batch still goes through the same calls as `lsh.query_batch` (K1 hashing + batched scan).
"""
from __future__ import annotations

import numpy as np

from . import lsh
from ._device import is_tensor, require_cuda, u64_table
from .ufuncs import hamming_top_n


class QueryStream:
    def __init__(self, hashes, projections, n: int = 10, depth: int = 2, flush_l2_bytes: int = 0):
        t = require_cuda()
        self.t = t
        self.H, _ = u64_table(hashes, "QueryStream", ndim=2)
        self.projections = projections
        self.n = int(n)
        self.depth = max(1, int(depth))
        dev = self.H.device
        with t.cuda.device(dev):
            self.s_in, self.s_run, self.s_out = (t.cuda.Stream(device=dev) for _ in range(3))
        self.slots = [dict(dev_in=None, pinned_in=None, pinned_out=None, ev_in=t.cuda.Event(), ev_run=t.cuda.Event(),
                           ev_out=t.cuda.Event(), busy=False, rows=None, q=0) for _ in range(self.depth)]
        self.count = 0
        # measurement aid (bench.py): evict L2 on the search stream before every batch
        self.flush = t.empty(flush_l2_bytes, dtype=t.uint8, device=dev) if flush_l2_bytes else None

    def submit(self, vectors) -> int:
        """Enqueue one batch (float32 [q, dims], NumPy or pinned CPU tensor); returns its ticket."""
        t = self.t
        slot = self.slots[self.count % self.depth]
        if slot["busy"]:
            slot["ev_out"].synchronize()          # its pinned buffers are about to be reused
        v = vectors if is_tensor(vectors) else t.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32))
        q, d = v.shape
        if slot["dev_in"] is None or slot["dev_in"].shape != (q, d):
            slot["dev_in"] = t.empty((q, d), dtype=t.float32, device=self.H.device)
            slot["pinned_out"] = t.empty((q, self.n), dtype=t.int64).pin_memory()
            slot["pinned_in"] = t.empty((q, d), dtype=t.float32).pin_memory()
        if not v.is_pinned():
            slot["pinned_in"].copy_(v)            # pageable source: stage through pinned memory
            v = slot["pinned_in"]
        with t.cuda.stream(self.s_in):
            slot["dev_in"].copy_(v, non_blocking=True)
            slot["ev_in"].record()
        with t.cuda.stream(self.s_run):
            self.s_run.wait_event(slot["ev_in"])
            if self.flush is not None:
                self.flush.fill_(self.count & 0xFF)
            sigs = lsh.hash_vectors(slot["dev_in"], self.projections)
            rows = hamming_top_n(self.H, sigs, n=self.n)
            slot["ev_run"].record()
        with t.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["ev_run"])
            rows.record_stream(self.s_out)
            slot["pinned_out"].copy_(rows.view(t.int64), non_blocking=True)
            slot["ev_out"].record()
        slot.update(busy=True, rows=rows, q=q)
        self.count += 1
        return self.count - 1

    def result(self, ticket: int) -> np.ndarray:
        """Block until batch `ticket` is back on the host; uint64 [q, n] view of its pinned buffer."""
        if not self.count - self.depth <= ticket < self.count:
            raise ValueError(f"ticket {ticket} is no longer (or not yet) in flight")
        slot = self.slots[ticket % self.depth]
        slot["ev_out"].synchronize()
        slot["busy"] = False
        return slot["pinned_out"].numpy().view(np.uint64)

    def map(self, batches):
        """Pipelined iteration: yields the ids of every batch, in order."""
        pending = []
        for b in batches:
            pending.append(self.submit(b))
            if len(pending) == self.depth:
                yield self.result(pending.pop(0))
        while pending:
            yield self.result(pending.pop(0))


class GraphedQueryBatch:
    """One fixed-shape batched query (hash the vectors, search, top-n) captured ONCE as a CUDA graph.

    A step of `lsh.query_batch` is ~20 dependent launches (K1T + fix-up + gated float64, then per phase
    of the tensor-core scan one scan and one select launch, then the gated repair launches) and no host
    synchronisation; for small per-GPU batches (a 10 000-query batch split over 8 GPUs) the gaps between
    those launches are a third of the step.  Replaying them as one graph removes the per-launch host cost
    and shortens the gaps.  Shapes, table and projections are fixed at capture; the vectors live in the
    static buffer `self.vectors` (write into it, or pass a tensor to __call__ to have it copied in).

        g = GraphedQueryBatch(hashes, projections, num_queries=10_000, dims=300, n=10)
        ids = g(vectors)            # uint64 [q, n] CUDA tensor, overwritten by the next call
    """

    def __init__(self, hashes, projections, num_queries: int, dims: int, n: int = 10, return_distances: bool = False,
                 vectors=None):
        t = require_cuda()
        self.t = t
        self.H, _ = u64_table(hashes, "GraphedQueryBatch", ndim=2)
        self.projections = projections
        self.n = int(n)
        self.return_distances = bool(return_distances)
        dev = self.H.device
        if vectors is not None:       # capture over the caller's own (resident, fixed-address) buffer
            if not (is_tensor(vectors) and vectors.is_cuda and vectors.dtype == t.float32 and vectors.is_contiguous()
                    and tuple(vectors.shape) == (num_queries, dims)):
                raise ValueError("vectors must be a contiguous float32 CUDA tensor [num_queries, dims]")
            self.vectors = vectors
        else:
            self.vectors = t.zeros((num_queries, dims), dtype=t.float32, device=dev)
        self.stream = t.cuda.Stream(device=dev)
        with t.cuda.device(dev):
            self.stream.wait_stream(t.cuda.current_stream())
            with t.cuda.stream(self.stream):          # warm-up outside the capture: allocator, kernel attributes
                for _ in range(2):
                    self._run()
            t.cuda.current_stream().wait_stream(self.stream)
            self.graph = t.cuda.CUDAGraph()
            with t.cuda.graph(self.graph, stream=self.stream):
                self.out = self._run()

    def _run(self):
        sigs = lsh.hash_vectors(self.vectors, self.projections)
        return hamming_top_n(self.H, sigs, n=self.n, return_distances=self.return_distances)

    def __call__(self, vectors=None):
        if vectors is not None and vectors is not self.vectors:
            self.vectors.copy_(vectors, non_blocking=True)
        self.graph.replay()
        return self.out
