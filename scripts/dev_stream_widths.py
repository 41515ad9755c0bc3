"""Development aid (GPU box): single-query streaming scan per width — kernel-only times from the
library's profiling events vs whole-call latency."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
for w in (1, 2, 4, 8, 10, 16, 32):
    H = torch.randint(-2**63, 2**63 - 1, (rows, w), dtype=torch.int64, device="cuda")
    for nq in (1, 2, 4, 8, 15):
        q = torch.randint(-2**63, 2**63 - 1, (nq, w), dtype=torch.int64, device="cuda")
        for _ in range(3):
            np_sims.hamming_top_n(H, q, 10)
        _lib.raw("nps_profile_enable")(1)
        sm, mm, tot = [], [], []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); np_sims.hamming_top_n(H, q, 10); e1.record(); torch.cuda.synchronize()
            s, m = _lib.profile_last(); sm.append(s); mm.append(m); tot.append(e0.elapsed_time(e1))
        _lib.raw("nps_profile_enable")(0)
        plan = _lib.last_scan_plan()
        gb = rows * w * 8 / 1e9
        print(f"W={w:2d} nq={nq}: scan {np.median(sm):.3f} ms ({gb / np.median(sm) * 1e3:.0f} GB/s, {rows / np.median(sm) / 1e6:.1f} Grows/s) merge {np.median(mm):.3f} ms, call {np.median(tot):.3f} ms; plan {plan}", flush=True)
    del H
