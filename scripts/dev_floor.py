"""Development aid (GPU box, NPS_KNOBS=1 build): which pipeline role sets the per-tile floor of mma_scan_kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch, np_sims
from np_sims import _lib
g = torch.Generator(device="cuda"); g.manual_seed(11)
H = torch.randint(-2**63, 2**63 - 1, (400_000, 10), dtype=torch.int64, device="cuda", generator=g)
Qs = torch.randint(-2**63, 2**63 - 1, (10_000, 10), dtype=torch.int64, device="cuda", generator=g)
_lib.set_scan_path(_lib.SCAN_MMA)
for pair in ("0", "1"):
    for dbg, name in ((0, "full"), (4, "no mma"), (8, "no epilogue filter"), (12, "no mma, no epilogue filter"), (2, "no expansion stores"),
                      (6, "no expansion, no mma"), (14, "no expansion, no mma, no epilogue filter"), (10, "no expansion, no epilogue filter")):
        os.environ["NPS_MMA_PAIR"] = pair; os.environ["NPS_MMA_QT1"] = "1"; os.environ["NPS_MMA_DEBUG"] = str(dbg)
        for _ in range(2):
            np_sims.hamming_top_n(H, Qs, 10)
        _lib.raw("nps_profile_enable")(1)
        np_sims.hamming_top_n(H, Qs, 10)
        ph = _lib.profile_last_phases()
        _lib.raw("nps_profile_enable")(0)
        t = ph[-1][1]
        print("pair=%s %-44s last phase %.3f ms = %.0f clk per (tile, query tile)" % (pair, name, t, t * 1e-3 * 1.965e9 * 148 / (1562 * 45)), flush=True)
