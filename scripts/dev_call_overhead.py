"""Development aid (GPU box): host-side cost of one canonical single-query hamming_top_n call."""
import os, sys, time, cProfile, pstats
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
H = torch.randint(-2**63, 2**63 - 1, (400_000, 10), dtype=torch.int64, device="cuda")
q = torch.randint(-2**63, 2**63 - 1, (10,), dtype=torch.int64, device="cuda")
for _ in range(50):
    np_sims.hamming_top_n(H, q, 10)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500):
    np_sims.hamming_top_n(H, q, 10)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("device tensors, no sync: host %.1f us per call; total with final sync %.1f us per call" % ((t1 - t0) / 500 * 1e6, (t2 - t0) / 500 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(500):
    np_sims.hamming_top_n(H, q, 10)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).strip_dirs().sort_stats("tottime").print_stats(14)
