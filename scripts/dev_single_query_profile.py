"""Development aid (GPU box): where the time of ONE reference-style call goes (lsh.query / hamming_top_10
through the searcher), at C2 size (400k x 640 bit) and on a 25M x 1024-bit table.  Run plain for host
timings; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel device times."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
from np_sims._device import attach
big = "--big" in sys.argv
rng = np.random.default_rng(0)
X = rng.standard_normal((400_000, 300), dtype=np.float32); X /= np.linalg.norm(X, axis=1, keepdims=True)
np.random.seed(0)
W = lsh.create_projections(640, 300)
H = lsh.index(X, W)
Q = X[:300].copy()
n = 20 if "--short" in sys.argv else 300
for v in Q[:20]:
    lsh.query(v, H, W)
t0 = time.perf_counter()
for v in Q[:n]:
    lsh.query(v, H, W)
print("lsh.query per call: %.1f us" % ((time.perf_counter() - t0) / n * 1e6))
s = lsh.index_one(Q[0], W)
for _ in range(5):
    np_sims.hamming_top_10(H, s)
t0 = time.perf_counter()
for v in Q[:n]:
    np_sims.hamming_top_10(H, s)
print("hamming_top_10 (searcher) per call: %.1f us" % ((time.perf_counter() - t0) / n * 1e6))
t0 = time.perf_counter()
for v in Q[:n]:
    np_sims.hamming_top_cand(H, s)
print("hamming_top_cand (searcher) per call: %.1f us" % ((time.perf_counter() - t0) / n * 1e6))
Hd = torch.from_numpy(np.asarray(H).view(np.int64)).cuda(); sd = torch.from_numpy(s.view(np.int64)).cuda()
for _ in range(5):
    np_sims.hamming_top_n(Hd, sd, 10, order="reference")
torch.cuda.synchronize(); t0 = time.perf_counter()
for v in Q[:n]:
    np_sims.hamming_top_n(Hd, sd, 10, order="reference")
torch.cuda.synchronize()
print("hamming_top_n(order=reference) device API per call: %.1f us" % ((time.perf_counter() - t0) / n * 1e6))
if big:
    Hb = torch.randint(-2**63, 2**63 - 1, (25_000_000, 16), dtype=torch.int64, device="cuda")
    qb = torch.randint(-2**63, 2**63 - 1, (16,), dtype=torch.int64, device="cuda")
    for k in (10, 1000):
        for _ in range(3):
            np_sims.hamming_top_n(Hb, qb, k, order="reference")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            np_sims.hamming_top_n(Hb, qb, k, order="reference")
        torch.cuda.synchronize()
        print("25M x 1024 bit, k=%d reference order: %.1f us per call" % (k, (time.perf_counter() - t0) / 5 * 1e6))
