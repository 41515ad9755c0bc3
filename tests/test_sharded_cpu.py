"""world_size-2 gloo test of the multi-GPU plumbing (np_sims.sharded) on CPU.

The CUDA kernels cannot run here, so the two injection points of ShardedIndex are filled with
oracle-backed stand-ins (TEST infrastructure): what is under test is the host logic — shard
bounds, row_base offsets, the all-gather layout handed to the merge, and that shard-then-merge
equals the single-shot search bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_local_keys(H, Q, k, row_base):
    out = np.full((len(Q), k), np.iinfo(np.uint64).max, dtype=np.uint64)
    for i, q in enumerate(Q):
        rows, dists = oracle.top_k_canonical(H, q, k)
        ok = rows != np.iinfo(np.uint64).max
        out[i, ok] = (dists[ok] << np.uint64(40)) | (rows[ok] + np.uint64(row_base))
    return torch.from_numpy(out.view(np.int64))


def _oracle_merge(gathered, k):
    g = gathered.numpy().view(np.uint64)                       # [G, q, k]
    merged = np.sort(np.transpose(g, (1, 0, 2)).reshape(g.shape[1], -1), axis=1)[:, :k]
    pad = merged == np.iinfo(np.uint64).max
    rows = np.where(pad, merged, merged & np.uint64((1 << 40) - 1))
    dists = np.where(pad, merged, merged >> np.uint64(40))
    return rows, dists


def _worker(rank, world, port, n, w, k, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "np-sims_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from np_sims.sharded import ShardedIndex, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    H = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64) & np.uint64(0xFFFF)
    Q = rng.integers(0, 2**64, size=(6, w), dtype=np.uint64) & np.uint64(0xFFFF)
    b = shard_bounds(n, world)
    assert b[0] == 0 and b[-1] == n and all(x % 2 == 0 for x in b[:-1])
    ix = ShardedIndex.from_global(H, rank, world, local_keys_fn=_oracle_local_keys, merge_fn=_oracle_merge)
    assert ix.row_base == b[rank] and len(ix.hashes) == b[rank + 1] - b[rank]
    rows, dists = ix.query(Q, k)
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), rows)
    np.save(os.path.join(out_dir, f"dists_{rank}.npy"), dists)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,w,k", [(1001, 3, 10), (37, 2, 100)])
def test_two_rank_shard_merge_equals_single_shot(tmp_path, n, w, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, w, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    H = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64) & np.uint64(0xFFFF)
    Q = rng.integers(0, 2**64, size=(6, w), dtype=np.uint64) & np.uint64(0xFFFF)
    r0, r1 = np.load(tmp_path / "rows_0.npy"), np.load(tmp_path / "rows_1.npy")
    d0, d1 = np.load(tmp_path / "dists_0.npy"), np.load(tmp_path / "dists_1.npy")
    assert np.array_equal(r0, r1) and np.array_equal(d0, d1)      # identical on every rank
    for i in range(len(Q)):
        want_r, want_d = oracle.top_k_canonical(H, Q[i], k)
        assert np.array_equal(r0[i], want_r) and np.array_equal(d0[i], want_d)


def test_shard_bounds_cover_and_align():
    from np_sims.sharded import shard_bounds
    for n in (0, 1, 2, 7, 1000, 100_000_001):
        for g in (1, 2, 4, 8):
            b = shard_bounds(n, g)
            assert b[0] == 0 and b[-1] == n and len(b) == g + 1
            assert all(b[i] <= b[i + 1] for i in range(g))
            assert all(x % 2 == 0 for x in b[:-1])


def _oracle_top_n(H, Q, k):
    rows = np.empty((len(Q), k), dtype=np.uint64)
    dists = np.empty((len(Q), k), dtype=np.uint64)
    for i, q in enumerate(Q):
        rows[i], dists[i] = oracle.top_k_canonical(H, q, k)
    return torch.from_numpy(rows.view(np.int64)), torch.from_numpy(dists.view(np.int64))


def _replicated_worker(rank, world, port, n, w, k, nq, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "np-sims_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from np_sims.sharded import ReplicatedIndex, query_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(321)
    H = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64) & np.uint64(0xFFFF)
    Q = rng.integers(0, 2**64, size=(nq, w), dtype=np.uint64) & np.uint64(0xFFFF)
    ix = ReplicatedIndex(H, top_n_fn=_oracle_top_n)
    b = query_bounds(nq, world)
    assert ix.my_slice(nq) == (b[rank], b[rank + 1])
    rows, dists = ix.query(Q, k)
    assert rows.shape == (nq, k)
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), rows.numpy().view(np.uint64))
    np.save(os.path.join(out_dir, f"dists_{rank}.npy"), dists.numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,w,k,nq", [(501, 3, 10, 7), (40, 2, 100, 6)])
def test_two_rank_query_sharded_equals_single_shot(tmp_path, n, w, k, nq):
    """Replicated table, ragged query slices (7 = 3 + 4): every rank ends with the full result."""
    world = 2
    mp.spawn(_replicated_worker, args=(world, _free_port(), n, w, k, nq, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(321)
    H = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64) & np.uint64(0xFFFF)
    Q = rng.integers(0, 2**64, size=(nq, w), dtype=np.uint64) & np.uint64(0xFFFF)
    r0, r1 = np.load(tmp_path / "rows_0.npy"), np.load(tmp_path / "rows_1.npy")
    d0, d1 = np.load(tmp_path / "dists_0.npy"), np.load(tmp_path / "dists_1.npy")
    assert np.array_equal(r0, r1) and np.array_equal(d0, d1)
    for i in range(nq):
        want_r, want_d = oracle.top_k_canonical(H, Q[i], k)
        assert np.array_equal(r0[i], want_r) and np.array_equal(d0[i], want_d)


def test_choose_sharding_and_query_bounds():
    from np_sims.sharded import choose_sharding, query_bounds
    assert choose_sharding(400_000, 10, 10_000, 1) == "rows"
    assert choose_sharding(400_000, 10, 10_000, 8) == "queries"        # C2: 32 MB table
    assert choose_sharding(100_000_000, 16, 100_000, 8) == "rows"      # C4: 12.8 GB table
    assert choose_sharding(400_000, 10, 100, 8) == "rows"              # too few queries per rank
    for q in (0, 1, 7, 10_000):
        for g in (1, 2, 3, 8):
            b = query_bounds(q, g)
            assert b[0] == 0 and b[-1] == q and all(b[i] <= b[i + 1] for i in range(g))


def test_recall_harness_helpers_cpu():
    """benchmark/recall.py pieces that need no GPU: synthetic data shape / norms, the split, recall@10."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "recall_harness", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmark", "recall.py"))
    rec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rec)
    x = rec.synthetic_vectors(500, 32, 10, 1)
    assert x.shape == (500, 32) and x.dtype == np.float32
    assert np.allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-5)
    q, ix = rec.test_train_split(x, 40, np.random.default_rng(0))
    assert q.shape == (40, 32) and ix.shape == (460, 32)
    ids = np.arange(20).reshape(2, 10)
    gt = np.array([np.arange(10), np.arange(5, 15)])
    assert rec.recall_at_10(ids, gt) == (1.0 + 0.5) / 2       # second query finds 5 of its 10 neighbours
