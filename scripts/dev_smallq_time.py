"""Development aid (GPU box): batched scan of a C2-size table with a SMALL query batch (one rank of an 8-GPU
query-sharded step) under the phase-schedule knobs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
q = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
H = torch.randint(-2**63, 2**63 - 1, (400_000, 10), dtype=torch.int64, device="cuda")
Q = torch.randint(-2**63, 2**63 - 1, (q, 10), dtype=torch.int64, device="cuda")
_lib.set_scan_path(_lib.SCAN_POPC); want = np_sims.hamming_top_n(H, Q, 10); _lib.set_scan_path(_lib.SCAN_AUTO)
for _ in range(5):
    got = np_sims.hamming_top_n(H, Q, 10)
torch.cuda.synchronize()
ms = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); got = np_sims.hamming_top_n(H, Q, 10); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ok = bool(torch.equal(got.view(torch.int64), want.view(torch.int64)))
from np_sims import ufuncs
print(f"q={q} GROWTH={os.environ.get('NPS_MMA_GROWTH')} LATE={os.environ.get('NPS_MMA_LATE')} CAP={os.environ.get('NPS_MMA_CAP')}: "
      f"{np.median(ms):.4f} ms, phases {_lib.last_mma_plan()['num_phases']}, equal {ok}, overflow {ufuncs.last_overflow_count()}")
_lib.raw("nps_profile_enable")(1)
np_sims.hamming_top_n(H, Q, 10); torch.cuda.synchronize()
print("  phases (tiles, ms):", [(t, round(m, 4)) for t, m in _lib.profile_last_phases()], "scan+select, repair:", _lib.profile_last())
_lib.raw("nps_profile_enable")(0)
