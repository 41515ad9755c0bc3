"""Development aid (GPU box): single-query scan of 64-bit signatures (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
w = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = torch.randint(-2**63, 2**63 - 1, (16_000_000, w), dtype=torch.int64, device="cuda")
q = torch.randint(-2**63, 2**63 - 1, (1, w), dtype=torch.int64, device="cuda")
for _ in range(3):
    np_sims.hamming_top_n(H, q, 10)
torch.cuda.synchronize()
print("ok")
