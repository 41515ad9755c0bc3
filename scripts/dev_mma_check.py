"""Development aid (GPU box): tensor-core scan vs XOR+POPC scan on a few shapes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib

def run(n, w, q, k, seed=0, planted=True):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    H = torch.randint(-2**63, 2**63 - 1, (n, w), dtype=torch.int64, device="cuda", generator=g)
    Q = torch.randint(-2**63, 2**63 - 1, (q, w), dtype=torch.int64, device="cuda", generator=g)
    if planted and n >= q:
        idx = torch.randperm(n, device="cuda", generator=g)[:q]
        Q = H[idx].clone()
        Q[:, 0] ^= 0x5
    _lib.set_scan_path(_lib.SCAN_POPC)
    r0, d0 = np_sims.hamming_top_n(H, Q, k, return_distances=True)
    torch.cuda.synchronize()
    _lib.set_scan_path(_lib.SCAN_MMA)
    t0 = time.perf_counter()
    r1, d1 = np_sims.hamming_top_n(H, Q, k, return_distances=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    path = _lib.raw("nps_last_scan_path")()
    plan = _lib.last_mma_plan() if path == _lib.SCAN_MMA else None
    _lib.set_scan_path(_lib.SCAN_AUTO)
    r0, d0, r1, d1 = (x.view(torch.int64).cpu().numpy() for x in (r0, d0, r1, d1))
    bad_r = int((r0 != r1).sum()); bad_d = int((d0 != d1).sum())
    print(f"n={n} w={w} q={q} k={k}: path={path} rows_bad={bad_r}/{r0.size} dists_bad={bad_d} t={1e3*(t1-t0):.2f}ms plan={plan}", flush=True)
    if bad_r or bad_d:
        qi = np.argwhere((r0 != r1).any(axis=1) | (d0 != d1).any(axis=1))[:3, 0]
        for i in qi:
            print("  q", i, "\n   popc rows", r0[i][:10], "d", d0[i][:10], "\n   mma  rows", r1[i][:10], "d", d1[i][:10])
    return bad_r + bad_d

if __name__ == "__main__":
    shapes = [(128, 2, 64, 10), (1000, 2, 64, 10), (5000, 10, 64, 10), (5000, 3, 100, 10), (70001, 10, 300, 10),
              (400000, 10, 1000, 10), (200000, 16, 512, 100), (50000, 5, 256, 1000), (100000, 32, 256, 10)]
    if len(sys.argv) > 1:
        shapes = shapes[: int(sys.argv[1])]
    bad = 0
    for s in shapes:
        bad += run(*s)
    print("TOTAL BAD", bad)
