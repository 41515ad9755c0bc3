"""Development aid (GPU box): batched scan of wide signatures, tensor cores vs POPC."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
for (n, w, nq, k) in [(2_000_000, 32, 1024, 10), (2_000_000, 40, 1024, 10), (2_000_000, 24, 1024, 10), (2_000_000, 23, 1024, 10), (2_000_000, 16, 1024, 10)]:
    H = torch.randint(-2**63, 2**63 - 1, (n, w), dtype=torch.int64, device="cuda")
    Q = torch.randint(-2**63, 2**63 - 1, (nq, w), dtype=torch.int64, device="cuda")
    res = {}
    for path in (_lib.SCAN_MMA, _lib.SCAN_POPC):
        prev = _lib.set_scan_path(path)
        try:
            np_sims.hamming_top_n(H, Q, k); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = np_sims.hamming_top_n(H, Q, k); e1.record(); torch.cuda.synchronize()
            res[path] = (e0.elapsed_time(e1), r)
            if path == _lib.SCAN_MMA:
                plan = _lib.last_mma_plan()
        finally:
            _lib.set_scan_path(prev)
    same = torch.equal(res[_lib.SCAN_MMA][1], res[_lib.SCAN_POPC][1])
    print(f"n={n} w={w} nq={nq}: mma {res[_lib.SCAN_MMA][0]:.2f} ms, popc {res[_lib.SCAN_POPC][0]:.2f} ms, same={same}, plan={plan}", flush=True)
    del H
