"""Development aid (GPU box): single / few-query hamming_top_n latency for k = 10..1000 (C5 corner)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
g = torch.Generator(device="cuda"); g.manual_seed(5)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
for bits in (64, 640, 1024):
    w = bits // 64
    H = torch.randint(-2**63, 2**63 - 1, (rows, w), dtype=torch.int64, device="cuda", generator=g)
    for batch in (1, 4, 16):
        Q = torch.randint(-2**63, 2**63 - 1, (batch, w), dtype=torch.int64, device="cuda", generator=g)
        line = []
        for k in (10, 100, 1000):
            prev = _lib.set_scan_path(_lib.SCAN_POPC)
            for _ in range(2):
                np_sims.hamming_top_n(H, Q, k)
            ms = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); np_sims.hamming_top_n(H, Q, k); e1.record(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            _lib.set_scan_path(prev)
            line.append("k=%d %.3f ms" % (k, float(np.median(ms))))
        print("bits=%d batch=%d: %s   (table %.2f GB -> %.3f ms at HBM peak)" % (bits, batch, "  ".join(line), rows * w * 8 / 1e9, rows * w * 8 / 6534.5e6))
    del H
