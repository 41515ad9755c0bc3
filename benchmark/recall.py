#!/usr/bin/env python
"""Recall@10 / QPS harness for the B200 LSH search path — the counterpart of the reference's
benchmark/glove.py (same flow: split off queries, exact-cosine ground truth, build or load the
index, run each algorithm, report mean recall@10 and QPS; glove.py:58-69,171-279) with the GPU
doing the ground truth and the search.

    python benchmark/recall.py --num_projections 128 640 2560 --algorithm lsh_batch lsh_pure_c
                               [--vectors my_vectors.npy] [--rows 400000] [--num_queries 1000]

Data: `--vectors file.npy` (float32 [N, D], e.g. the GloVe matrix the reference caches as
data/glove.840B.300d_vectors.npy) or, by default, a synthetic GloVe-shaped set — unit vectors drawn
around `--clusters` random centres so that nearest neighbours mean something (there is no network
here to fetch GloVe).  Index files use the reference's names and format (benchmark/glove.py:262-279,
data_dir.py:10-17): data/hashes_{N}_{B}_{D}_{seed}.npy = np.save of uint64 [N, B/64],
data/projs_{N}_{B}_{D}_{seed}.npy = float64 [B, D] — files written by either implementation load in
the other.

Algorithms: lsh_batch (lsh.query_batch: the whole query set in one call — the B200 throughput
path), lsh_pure_c (lsh.query per vector, the reference's `lsh_pure_c`), lsh_rerank_cand
(lsh.query_two_phase per vector, needs --num_fronts), lsh_rerank_batch (lsh.query_two_phase_batch).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "np-sims_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA_PATH = "data/"


def synthetic_vectors(rows, dims, clusters, seed, spread=0.36):
    """Unit vectors = random unit centre + Gaussian noise of norm ~`spread` (cos ~0.9 inside a cluster)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((clusters, dims)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    x = centres[rng.integers(0, clusters, size=rows)] + np.float32(spread / np.sqrt(dims)) * rng.standard_normal((rows, dims), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def test_train_split(vectors, num_queries, rng):
    """Split off num_queries rows as queries and remove them from the index set (glove.py:282-291)."""
    q = rng.choice(len(vectors), size=num_queries, replace=False)
    return vectors[q], np.delete(vectors, q, axis=0)


def most_similar_cos(index_vectors, query_vectors, n=10, chunk=2048):
    """Exact ground truth (glove.py:58-69) on the GPU: top-n by dot product, [q, n] ids."""
    import torch
    X = torch.from_numpy(index_vectors).cuda()
    out = []
    for i in range(0, len(query_vectors), chunk):
        Q = torch.from_numpy(query_vectors[i:i + chunk]).cuda()
        out.append(torch.topk(Q @ X.T, n, dim=1).indices.cpu().numpy())
    return np.concatenate(out)


def load_or_build_index(index_vectors, num_projections, seed):
    """glove.py:262-279 with the same file names and formats; the build runs on the GPU (K1T)."""
    import np_sims.lsh as lsh
    os.makedirs(DATA_PATH, exist_ok=True)
    dims = index_vectors.shape[1]
    suffix = f"{len(index_vectors)}_{num_projections}_{dims}_{seed}"
    hf, pf = os.path.join(DATA_PATH, f"hashes_{suffix}.npy"), os.path.join(DATA_PATH, f"projs_{suffix}.npy")
    if os.path.exists(hf) and os.path.exists(pf):
        print(f"loaded from {hf} and {pf}")
        return np.load(hf), np.load(pf), 0.0
    if seed is not None:
        np.random.seed(seed)
    projs = lsh.create_projections(num_projections, dims)
    t0 = time.perf_counter()
    hashes = lsh.index(index_vectors, projs)
    build_s = time.perf_counter() - t0
    np.save(hf, np.asarray(hashes))
    np.save(pf, np.asarray(projs))
    print(f"saved {hf} and {pf}")
    return hashes, projs, build_s


def run_algorithm(name, query_vectors, hashes, projs, num_fronts):
    """-> (ids [q, 10], seconds)"""
    import torch
    import np_sims.lsh as lsh
    front = lsh.front_hashes(hashes, num_fronts) if num_fronts else None
    if name in ("lsh_batch", "lsh_rerank_batch"):
        fn = (lambda: lsh.query_batch(query_vectors, hashes, projs, n=10)) if name == "lsh_batch" else \
             (lambda: lsh.query_two_phase_batch(query_vectors, front, hashes, projs, n=10))
        fn()                                    # warm-up (uploads, plan caches)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids = fn()
        torch.cuda.synchronize()
        return np.asarray(ids), time.perf_counter() - t0
    one = {"lsh_pure_c": lambda v: lsh.query(v, hashes, projs),
           "lsh_pure_python": lambda v: lsh.query_pure_python(v, hashes, projs),
           "lsh_rerank_cand": lambda v: lsh.query_two_phase(v, front, hashes, projs)}[name]
    one(query_vectors[0])
    t0 = time.perf_counter()
    ids = [np.asarray(one(v)[0])[:10] for v in query_vectors]
    return np.stack(ids), time.perf_counter() - t0


def recall_at_10(ids, gt):
    return float(np.mean([len(set(a.tolist()) & set(b.tolist())) / len(b) for a, b in zip(ids, gt)]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algorithm", nargs="+", default=["lsh_batch", "lsh_pure_c"])
    ap.add_argument("--num_projections", nargs="+", type=int, default=[64, 128, 256, 640, 1280, 2560])
    ap.add_argument("--num_fronts", nargs="+", type=int, default=[None])
    ap.add_argument("--num_queries", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--vectors", default=None, help=".npy float32 [N, D]; default: synthetic clustered unit vectors")
    ap.add_argument("--rows", type=int, default=400_000)
    ap.add_argument("--dims", type=int, default=300)
    ap.add_argument("--clusters", type=int, default=4000)
    ap.add_argument("--per_query_limit", type=int, default=200, help="queries run by the one-vector-at-a-time algorithms")
    ap.add_argument("--out", default=None, help="write the result table as JSON")
    args = ap.parse_args()

    rng = np.random.default_rng(args.seed)
    vectors = np.load(args.vectors).astype(np.float32) if args.vectors else \
        synthetic_vectors(args.rows + args.num_queries, args.dims, args.clusters, args.seed)
    query_vectors, index_vectors = test_train_split(vectors, args.num_queries, rng)
    gt = most_similar_cos(index_vectors, query_vectors)
    print(f"{len(index_vectors)} index vectors, {len(query_vectors)} queries, dims {index_vectors.shape[1]}")
    table = []
    for b in args.num_projections:
        hashes, projs, build_s = load_or_build_index(index_vectors, b, args.seed)
        for algo in args.algorithm:
            for front in (args.num_fronts if "rerank" in algo else [None]):
                if "rerank" in algo and front is None:
                    continue
                batched = algo in ("lsh_batch", "lsh_rerank_batch")
                qv = query_vectors if batched else query_vectors[:args.per_query_limit]
                ids, secs = run_algorithm(algo, qv, hashes, projs, front)
                row = {"algorithm": algo, "num_projections": b, "num_fronts": front, "queries": len(qv),
                       "recall_at_10": recall_at_10(ids, gt[:len(qv)]), "qps": len(qv) / secs, "index_build_s": build_s}
                table.append(row)
                print(f"{algo:17s} {b:5d} projections" + (f" fronts {front}" if front else "") +
                      f" | Mean recall@10: {row['recall_at_10']:.4f} | QPS: {row['qps']:.1f}")
    if args.out:
        json.dump({"rows": len(index_vectors), "dims": int(index_vectors.shape[1]),
                   "data": args.vectors or f"synthetic, {args.clusters} clusters, seed {args.seed}", "results": table},
                  open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
