"""Development aid (GPU box): batched two-phase search vs the exact full-width batched scan on C2."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
g = torch.Generator(device="cuda"); g.manual_seed(0)
X = torch.randn((400_000, 300), device="cuda", generator=g); X /= X.norm(dim=1, keepdim=True)
Qv = torch.randn((10_000, 300), device="cuda", generator=g); Qv /= Qv.norm(dim=1, keepdim=True)
np.random.seed(0)
W = lsh.create_projections(640, 300)
H = lsh.hash_vectors(X, W)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return np.median(ms)
print("exact full-width batch:", t(lambda: lsh.query_batch(Qv, H, W, n=10)), "ms")
for fw, nc in ((1, 1000), (2, 1000), (2, 100), (4, 100)):
    F = H[:, :fw].contiguous()
    print(f"two-phase front={fw*64} bits, {nc} candidates:", t(lambda: lsh.query_two_phase_batch(Qv, F, H, W, n=10, num_candidates=nc)), "ms")
    a = lsh.query_two_phase_batch(Qv[:256], F, H, W, n=10, num_candidates=nc)
    b = lsh.query_batch(Qv[:256], H, W, n=10)
    inter = np.mean([len(set(x.tolist()) & set(y.tolist())) / 10 for x, y in zip(a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64))])
    print(f"   overlap with the exact top-10: {inter:.3f}")
