"""Development aid (GPU box, NPS_KNOBS=1 build): C2-shaped tensor-core scan under different knob settings
(CTA pairs on/off, A-ring grouping, debug switches that remove one pipeline role at a time)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib
g = torch.Generator(device="cuda"); g.manual_seed(11)
N, W, Q, k = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (400_000, 10, 10_000, 10)))
H = torch.randint(-2**63, 2**63 - 1, (N, W), dtype=torch.int64, device="cuda", generator=g)
Qs = torch.randint(-2**63, 2**63 - 1, (Q, W), dtype=torch.int64, device="cuda", generator=g)
_lib.set_scan_path(_lib.SCAN_MMA)
configs = [
    ("single", {"NPS_MMA_PAIR": "0"}),
    ("pair", {"NPS_MMA_PAIR": "1"}),
    ("single groups 1,3", {"NPS_MMA_PAIR": "0", "NPS_MMA_GROUPS": "1,3"}),
    ("single groups 1,3 no-mma", {"NPS_MMA_PAIR": "0", "NPS_MMA_GROUPS": "1,3", "NPS_MMA_DEBUG": "4"}),
    ("pair groups 1,3 qt1", {"NPS_MMA_PAIR": "1", "NPS_MMA_QT1": "1", "NPS_MMA_GROUPS": "1,3"}),
    ("pair groups 1,4 qt1", {"NPS_MMA_PAIR": "1", "NPS_MMA_QT1": "1", "NPS_MMA_GROUPS": "1,4"}),
    ("pair one-qtile", {"NPS_MMA_PAIR": "1", "NPS_MMA_QT1": "1"}),
    ("pair no-mma(4) qt2", {"NPS_MMA_PAIR": "1", "NPS_MMA_DEBUG": "4"}),
    ("pair pstages 2", {"NPS_MMA_PAIR": "1", "NPS_MMA_PSTAGES": "2"}),
    ("pair pstages 4", {"NPS_MMA_PAIR": "1", "NPS_MMA_PSTAGES": "4"}),
    ("pair pstages 4 no-mma", {"NPS_MMA_PAIR": "1", "NPS_MMA_PSTAGES": "4", "NPS_MMA_DEBUG": "4"}),
    ("pair pstages 6 groups 2,2", {"NPS_MMA_PAIR": "1", "NPS_MMA_GROUPS": "2,2"}),
    ("pair groups 2,2", {"NPS_MMA_PAIR": "1", "NPS_MMA_GROUPS": "2,2"}),
    ("pair groups 1,2", {"NPS_MMA_PAIR": "1", "NPS_MMA_GROUPS": "1,2"}),
    ("pair groups 1,1", {"NPS_MMA_PAIR": "1", "NPS_MMA_GROUPS": "1,1"}),
    ("pair no-epilogue-filter(1)", {"NPS_MMA_PAIR": "1", "NPS_MMA_DEBUG": "1"}),
    ("pair no-expansion(2)", {"NPS_MMA_PAIR": "1", "NPS_MMA_DEBUG": "2"}),
    ("pair no-mma(4)", {"NPS_MMA_PAIR": "1", "NPS_MMA_DEBUG": "4"}),
    ("pair no-epilogue(8)", {"NPS_MMA_PAIR": "1", "NPS_MMA_DEBUG": "8"}),
    ("single no-expansion(2)", {"NPS_MMA_PAIR": "0", "NPS_MMA_DEBUG": "2"}),
    ("single no-mma(4)", {"NPS_MMA_PAIR": "0", "NPS_MMA_DEBUG": "4"}),
    ("single no-epilogue(8)", {"NPS_MMA_PAIR": "0", "NPS_MMA_DEBUG": "8"}),
]
knobs = ("NPS_MMA_PAIR", "NPS_MMA_GROUPS", "NPS_MMA_DEBUG", "NPS_MMA_PSTAGES", "NPS_MMA_QT1")
for name, env in configs:
    for kk in knobs:
        os.environ.pop(kk, None)
    os.environ.update(env)
    try:
        for _ in range(2):
            np_sims.hamming_top_n(H, Qs, k)
        torch.cuda.synchronize()
        _lib.raw("nps_profile_enable")(1)
        runs = []
        for _ in range(3):
            np_sims.hamming_top_n(H, Qs, k)
            runs.append((_lib.profile_last()[0], _lib.profile_last_phases()))
        _lib.raw("nps_profile_enable")(0)
        tot = float(np.median([r[0] for r in runs]))
        ph = runs[-1][1]
        plan = _lib.last_mma_plan()
        print("%-28s total %.3f ms  phases %s  smem %d a_stages %d" % (name, tot, [round(m, 3) for _, m in ph], plan["smem_bytes"], plan["a_stages"]), flush=True)
        if "--check" in sys.argv and "DEBUG" not in "".join(env):
            got = np_sims.hamming_top_n(H, Qs, k)
            os.environ["NPS_MMA_PAIR"] = "0"; os.environ.pop("NPS_MMA_PSTAGES", None); os.environ.pop("NPS_MMA_GROUPS", None); os.environ.pop("NPS_MMA_QT1", None)
            want = np_sims.hamming_top_n(H, Qs, k)
            print("    equal to single:", bool(torch.equal(got.view(torch.int64), want.view(torch.int64))), flush=True)
    except Exception as e:
        print(name, "FAILED", e, flush=True)
