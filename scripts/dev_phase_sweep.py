"""Development aid (GPU box, NPS_KNOBS=1 build): phase schedule of the tensor-core scan (candidate capacity, growth
per phase, number of doubling phases at the end) against step time, for a full batch and for an eighth of it."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch, np_sims
from np_sims import _lib
g = torch.Generator(device="cuda"); g.manual_seed(11)
H = torch.randint(-2**63, 2**63 - 1, (400_000, 10), dtype=torch.int64, device="cuda", generator=g)
_lib.set_scan_path(_lib.SCAN_MMA)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for Q in (1250, 2500):
    Qs = torch.randint(-2**63, 2**63 - 1, (Q, 10), dtype=torch.int64, device="cuda", generator=g)
    want = None
    for cap, growth, late in ((512, 8, 3), (1024, 16, 2), (2048, 16, 2), (1024, 8, 2), (2048, 8, 2), (512, 8, 2)):
        for kk in ("NPS_MMA_CAP", "NPS_MMA_GROWTH", "NPS_MMA_LATE"):
            os.environ.pop(kk, None)
        if cap:
            os.environ.update(NPS_MMA_CAP=str(cap), NPS_MMA_GROWTH=str(growth), NPS_MMA_LATE=str(late))
        for _ in range(3):
            out = np_sims.hamming_top_n(H, Qs, 10)
        ms = []
        for i in range(7):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = np_sims.hamming_top_n(H, Qs, 10); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        from np_sims.ufuncs import last_overflow_count
        ov = last_overflow_count()
        if want is None:
            want = out.clone()
        same = bool(torch.equal(out.view(torch.int64), want.view(torch.int64)))
        print("Q=%5d cap=%4d growth=%3d late=%d: %.3f ms  phases=%d overflow=%d same=%s" % (Q, cap, growth, late, float(np.median(ms)), _lib.last_mma_plan()["num_phases"], ov, same), flush=True)
