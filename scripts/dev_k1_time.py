"""Development aid (GPU box): time K1T on the C2 index shape and a C3 slice; run under ncu
(--metrics gpu__time_duration.sum) for the per-kernel split."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims.lsh as lsh

def run(n, d, b, reps=3):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    X = torch.randn((n, d), device="cuda", generator=g)
    X /= X.norm(dim=1, keepdim=True)
    np.random.seed(0)
    W = lsh.create_projections(b, d)
    lsh.hash_vectors(X, W); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lsh.hash_vectors(X, W); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    st = lsh.hash_stats()
    flop = 2.0 * n * d * b
    print(f"n={n} d={d} b={b}: {min(ms):.3f} ms best ({flop / min(ms) / 1e9:.1f} TFLOP/s algorithmic, "
          f"{n * d * 4 / min(ms) / 1e6:.0f} GB/s of X), recomputed {st.get('recomputed')} of {st.get('pairs')}", flush=True)

if __name__ == "__main__":
    run(10_000, 300, 640)
    run(400_000, 300, 640)
    run(1_000_000, 768, 1024)
