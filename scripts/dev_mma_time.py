"""Development aid (GPU box): time the tensor-core scan on the C2 shape with CUDA events."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib

n, w, q, k = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (400000, 10, 10000, 10)))
g = torch.Generator(device="cuda"); g.manual_seed(0)
H = torch.randint(-2**63, 2**63 - 1, (n, w), dtype=torch.int64, device="cuda", generator=g)
Q = torch.randint(-2**63, 2**63 - 1, (q, w), dtype=torch.int64, device="cuda", generator=g)
_lib.set_scan_path(_lib.SCAN_MMA)
for _ in range(3):
    np_sims.hamming_top_n(H, Q, k)
torch.cuda.synchronize()
_lib.raw("nps_profile_enable")(1)
tot, last = [], []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); np_sims.hamming_top_n(H, Q, k); e1.record(); torch.cuda.synchronize()
    tot.append(e0.elapsed_time(e1)); last.append(_lib.profile_last()[0])
print(f"debug={os.environ.get('NPS_MMA_DEBUG')} n={n} w={w} q={q} k={k} total_ms={np.mean(tot):.3f} last_phase_scan_ms={np.mean(last):.3f} plan={_lib.last_mma_plan()}")
