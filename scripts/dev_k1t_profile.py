"""K1T alone for ncu: 2M x 768 fp32 -> 1024-bit signatures (C3 shape, 1/5 of the rows), one warm call + one profiled.

    ncu --set full --clock-control none -k regex:project_ -s 3 -c 3 -o /tmp/k1t python scripts/dev_k1t_profile.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims.lsh as lsh

g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.empty((2_000_000, 768), dtype=torch.float32, device="cuda").normal_(generator=g)
np.random.seed(0)
W = lsh.create_projections(1024, 768)
for _ in range(2):
    lsh.hash_vectors(X, W)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lsh.hash_vectors(X, W); e1.record(); torch.cuda.synchronize()
print(f"2M x 768 -> 1024 bit: {e0.elapsed_time(e1):.3f} ms")
