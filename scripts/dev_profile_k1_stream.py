"""Development aid (GPU box): the workloads profiled for profiles/r01_stream_w10.txt and
profiles/r01_project_tc_v2.txt — single-query streaming scan (16M x 640-bit) and K1T (400k x 300 -> 640)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
H = torch.randint(-2**63, 2**63 - 1, (16_000_000, 10), dtype=torch.int64, device="cuda")
q = torch.randint(-2**63, 2**63 - 1, (1, 10), dtype=torch.int64, device="cuda")
for _ in range(3):
    np_sims.hamming_top_n(H, q, 10)
torch.cuda.synchronize()
del H
g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.randn((400_000, 300), device="cuda", generator=g)
X /= X.norm(dim=1, keepdim=True)
np.random.seed(0)
W = lsh.create_projections(640, 300)
for _ in range(2):
    lsh.hash_vectors(X, W)
torch.cuda.synchronize()
print("ok")
