"""Development aid (GPU box): tensor-core hashing vs the float64 kernel."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims.lsh as lsh

def run(n, d, b, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d), dtype=np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    np.random.seed(0)
    W = lsh.create_projections(b, d)
    Xd = torch.from_numpy(X).cuda()
    ref = lsh.hash_vectors(Xd, W, exact_only=True)
    torch.cuda.synchronize()
    got = lsh.hash_vectors(Xd, W)
    torch.cuda.synchronize()
    st = lsh.hash_stats()
    diff = (ref ^ got)
    nbad = int(sum(bin(int(v) & (2**64 - 1)).count("1") for v in diff.flatten().tolist())) if n * b <= 2_000_000 else int((diff != 0).sum().item())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); lsh.hash_vectors(Xd, W, exact_only=True); ev[1].record()
    ev[2].record(); lsh.hash_vectors(Xd, W); ev[3].record(); torch.cuda.synchronize()
    print(f"n={n} d={d} b={b}: differing bits/words={nbad} of {n*b}; stats={st} frac={st['recomputed']/max(st['pairs'],1):.4f}; f64 {ev[0].elapsed_time(ev[1]):.3f} ms, tc {ev[2].elapsed_time(ev[3]):.3f} ms", flush=True)
    return nbad

if __name__ == "__main__":
    bad = 0
    for s in [(128, 32, 64), (1000, 300, 640), (5000, 300, 640), (10000, 300, 640), (4097, 768, 1024), (3000, 100, 128), (400000, 300, 640)]:
        bad += run(*s)
    print("TOTAL BAD", bad)
