"""Development aid (GPU box): where the time of ONE reference-style lsh.query call goes."""
import os, sys, time, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
rng = np.random.default_rng(0)
X = rng.standard_normal((400_000, 300), dtype=np.float32); X /= np.linalg.norm(X, axis=1, keepdims=True)
np.random.seed(0)
W = lsh.create_projections(640, 300)
H = lsh.index(X, W)
Q = X[:300].copy()
for v in Q[:20]:
    lsh.query(v, H, W)
t0 = time.perf_counter()
for v in Q:
    lsh.query(v, H, W)
print("lsh.query per call: %.3f ms" % ((time.perf_counter() - t0) / len(Q) * 1e3))
t0 = time.perf_counter()
for v in Q:
    s = lsh.index_one(v, W)
print("  index_one: %.3f ms" % ((time.perf_counter() - t0) / len(Q) * 1e3))
t0 = time.perf_counter()
for v in Q:
    np_sims.hamming_top_10(H, s)
print("  hamming_top_10: %.3f ms" % ((time.perf_counter() - t0) / len(Q) * 1e3))
t0 = time.perf_counter()
for v in Q:
    np_sims.hamming_top_n(H, s, 10)
print("  hamming_top_n canonical: %.3f ms" % ((time.perf_counter() - t0) / len(Q) * 1e3))
pr = cProfile.Profile(); pr.enable()
for v in Q:
    lsh.query(v, H, W)
pr.disable()
pstats.Stats(pr).strip_dirs().sort_stats("cumulative").print_stats(28)
