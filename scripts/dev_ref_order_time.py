"""Development aid (GPU box): reference-order single query (K5 + replay) timing pieces."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
H = torch.randint(-2**63, 2**63 - 1, (400_000, 10), dtype=torch.int64, device="cuda")
q = torch.randint(-2**63, 2**63 - 1, (10,), dtype=torch.int64, device="cuda")
for k, fn in (("top_10", np_sims.hamming_top_10), ("top_cand", np_sims.hamming_top_cand)):
    for _ in range(3):
        fn(H, q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        fn(H, q)
    torch.cuda.synchronize()
    print(k, "%.3f ms per call" % ((time.perf_counter() - t0) / 20 * 1e3), flush=True)
