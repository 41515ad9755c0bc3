"""Development aid (GPU box): small invocations of every kernel family, for compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
from np_sims import _lib
rng = np.random.default_rng(0)
X = rng.standard_normal((3001, 300)).astype(np.float32)
np.random.seed(0)
W = lsh.create_projections(640, 300)
H = lsh.index(X, W)                                   # K1T + fix-up + gated f64
Hd = torch.from_numpy(np.asarray(H).view(np.int64)).cuda()
q = lsh.hash_vectors(torch.from_numpy(X[:130]).cuda(), W)
for path in (_lib.SCAN_MMA, _lib.SCAN_POPC):
    _lib.set_scan_path(path)
    for k in (10, 100):
        np_sims.hamming_top_n(Hd, q, k)                # K2T phases + selects + gated repair / K2 + K3
_lib.set_scan_path(_lib.SCAN_AUTO)
np_sims.hamming_top_n(Hd, q[:3], 10)
np_sims.hamming_top_10(Hd, q[0])                      # K5 + replay
np_sims.hamming_top_cand(Hd, q[0])
np_sims.hamming_c(Hd, q[0])
c = np_sims.hamming_top_n(Hd[:, :2].contiguous(), q[:, :2].contiguous(), 100)
np_sims.hamming_rerank(Hd, q, c, 10)                  # K6
w = torch.randint(-2**63, 2**63 - 1, (2000, 33), dtype=torch.int64, device="cuda")
_lib.set_scan_path(_lib.SCAN_MMA)
np_sims.hamming_top_n(w, w[:70].contiguous(), 10)     # wide signatures, three... two threshold slices
_lib.set_scan_path(_lib.SCAN_AUTO)
torch.cuda.synchronize()
print("sanitize workload done")
