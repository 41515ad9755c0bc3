"""Development aid (GPU box): run the tensor-core scan a few times on one shape (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
n, w, q, k, reps = (int(x) for x in (sys.argv[1:6] if len(sys.argv) > 5 else (400000, 10, 10000, 10, 1)))
g = torch.Generator(device="cuda"); g.manual_seed(0)
H = torch.randint(-2**63, 2**63 - 1, (n, w), dtype=torch.int64, device="cuda", generator=g)
Q = torch.randint(-2**63, 2**63 - 1, (q, w), dtype=torch.int64, device="cuda", generator=g)
_lib.set_scan_path(_lib.SCAN_MMA)
for _ in range(reps):
    np_sims.hamming_top_n(H, Q, k)
torch.cuda.synchronize()
print("done", _lib.last_mma_plan())
