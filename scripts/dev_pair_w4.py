import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/np-sims_b200"]
import numpy as np, torch, np_sims
from np_sims import _lib
from np_sims.ufuncs import last_overflow_count
from oracle import oracle
rng = np.random.default_rng(11)
_lib.set_scan_path(_lib.SCAN_MMA)
for (N, W, Q, k) in ((4_000_000, 4, 10_000, 10), (1_000_000, 4, 2_000, 10), (300_000, 4, 600, 10), (4_000_000, 5, 10_000, 10)):
    h = rng.integers(0, 2**64, size=(N, W), dtype=np.uint64)
    q = rng.integers(0, 2**64, size=(Q, W), dtype=np.uint64)
    H = torch.from_numpy(h.view(np.int64)).cuda(); Qs = torch.from_numpy(q.view(np.int64)).cuda()
    sel = rng.choice(Q, size=24, replace=False)
    want = [oracle.top_k_canonical(h, q[i], k)[0] for i in sel]
    for pair in (0, 1):
        for ps in ("2", "8"):
            os.environ["NPS_MMA_PAIR"] = str(pair); os.environ["NPS_MMA_PSTAGES"] = ps
            out = np_sims.hamming_top_n(H, Qs, k)
            ov = last_overflow_count()
            out = out.view(torch.int64).cpu().numpy().view(np.uint64)
            bad = sum(not np.array_equal(out[i], w) for i, w in zip(sel, want))
            print(f"N={N} W={W} Q={Q} pair={pair} pstages={ps}: wrong {bad}/24 overflowed_queries={ov} plan={_lib.last_mma_plan()}", flush=True)
