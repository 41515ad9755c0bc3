"""Development aid (GPU box): single-query scan latency vs k."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
H = torch.randint(-2**63, 2**63 - 1, (16_000_000, 10), dtype=torch.int64, device="cuda")
for nq in (1, 2):
    q = torch.randint(-2**63, 2**63 - 1, (nq, 10), dtype=torch.int64, device="cuda")
    for k in (1, 10, 32, 33, 100, 128, 129, 1000):
        for _ in range(3):
            np_sims.hamming_top_n(H, q, k)
        _lib.raw("nps_profile_enable")(1)
        sm = []
        for _ in range(5):
            np_sims.hamming_top_n(H, q, k); torch.cuda.synchronize(); sm.append(_lib.profile_last())
        _lib.raw("nps_profile_enable")(0)
        print(f"nq={nq} k={k}: scan {np.median([a for a, b in sm]):.3f} ms merge {np.median([b for a, b in sm]):.3f} ms", flush=True)
