"""Development aid (GPU box): C4-shaped shard (1024-bit signatures, k=100): tensor-core scan vs
XOR+POPC scan on a query sample, and throughput."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib

n, w, q, k = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (12_500_000, 16, 100_000, 100)))
g = torch.Generator(device="cuda"); g.manual_seed(3)
H = torch.randint(-2**63, 2**63 - 1, (n, w), dtype=torch.int64, device="cuda", generator=g)
g.manual_seed(4)
Q = torch.randint(-2**63, 2**63 - 1, (q, w), dtype=torch.int64, device="cuda", generator=g)
# planted neighbours for half of the queries: a table row with r in [0, 64] flipped bits
idx = torch.randint(0, n, (q // 2,), device="cuda", generator=g)
Q[: q // 2] = H[idx]
flip = torch.randint(-2**63, 2**63 - 1, (q // 2,), dtype=torch.int64, device="cuda", generator=g) & 0x0101010101010101
Q[: q // 2, 0] ^= flip
_lib.set_scan_path(_lib.SCAN_AUTO)
for _ in range(1):
    r1, d1 = np_sims.hamming_top_n(H, Q, k, return_distances=True)
torch.cuda.synchronize()
assert _lib.raw("nps_last_scan_path")() == _lib.SCAN_MMA
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.raw("nps_profile_enable")(1)
e0.record(); r1, d1 = np_sims.hamming_top_n(H, Q, k, return_distances=True); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
phases = _lib.profile_last_phases(); _lib.raw("nps_profile_enable")(0)
flop = 2.0 * n * q * w * 64
print(f"n={n} w={w} q={q} k={k}: {ms:.1f} ms -> {q/ms*1e3:.0f} q/s, {flop/ms/1e9:.0f} TFLOP/s; plan={_lib.last_mma_plan()}")
print("phases (tiles, ms):", [(t, round(m, 2)) for t, m in phases])
sample = torch.cat([torch.arange(0, 32, device="cuda"), torch.arange(q - 31, q, device="cuda"), torch.randint(0, q, (min(q, 2000),), device="cuda", generator=g)])
_lib.set_scan_path(_lib.SCAN_POPC)
r0, d0 = np_sims.hamming_top_n(H, Q[sample].contiguous(), k, return_distances=True)
torch.cuda.synchronize()
_lib.set_scan_path(_lib.SCAN_AUTO)
ok = torch.equal(r0.view(torch.int64), r1.view(torch.int64)[sample]) and torch.equal(d0.view(torch.int64), d1.view(torch.int64)[sample])
print("tensor-core == popc on", len(sample), "sampled queries:", ok, " best distances of planted query 0:", d1.view(torch.int64)[0, :3].tolist())
sys.exit(0 if ok else 1)
