import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/np-sims_b200"]
import numpy as np, torch, np_sims
from np_sims import _lib
g = torch.Generator(device="cuda"); g.manual_seed(11)
_lib.set_scan_path(_lib.SCAN_MMA)
for (N, W, Q, k) in ((3_000_000, 16, 20_000, 100), (1_000_000, 32, 4096, 10), (4_000_000, 4, 10_000, 10), (400_000, 10, 1250, 10)):
    H = torch.randint(-2**63, 2**63 - 1, (N, W), dtype=torch.int64, device="cuda", generator=g)
    Qs = torch.randint(-2**63, 2**63 - 1, (Q, W), dtype=torch.int64, device="cuda", generator=g)
    res = {}
    for pair in (0, 1):
        os.environ["NPS_MMA_PAIR"] = str(pair)
        for _ in range(2):
            out = np_sims.hamming_top_n(H, Qs, k)
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = np_sims.hamming_top_n(H, Qs, k); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        res[pair] = out
        pl = _lib.last_mma_plan()
        print(f"N={N} W={W} Q={Q} k={k} pair={pair}: {np.median(ms):.3f} ms  tflops={2.0*N*Q*W*64/np.median(ms)/1e9:.0f}  plan n_tile={pl['n_tile']} a_stages={pl['a_stages']} smem={pl['smem_bytes']} phases={pl['num_phases']}", flush=True)
    print("   equal:", bool(torch.equal(res[0].view(torch.int64), res[1].view(torch.int64))))
    del H, Qs
