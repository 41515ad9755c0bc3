"""Development aid (GPU box): K1T operand encodings (BF16 x 2 vs TF32 x 2) over table sizes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims.lsh as lsh
from np_sims import _lib
for (d, b) in ((300, 640), (768, 1024)):
    np.random.seed(0)
    W = lsh.create_projections(b, d)
    for n in (1000, 10_000, 30_000, 100_000, 400_000, 1_000_000):
        X = torch.randn((n, d), device="cuda"); X /= X.norm(dim=1, keepdim=True)
        out = {}
        for path, name in ((_lib.HASH_TF32, "tf32"), (_lib.HASH_BF16, "bf16")):
            _lib.call("nps_set_hash_path", path)
            lsh.hash_vectors(X, W); torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); lsh.hash_vectors(X, W); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
            out[name] = min(ms)
        _lib.call("nps_set_hash_path", _lib.HASH_AUTO)
        print(f"d={d} b={b} n={n}: tf32 {out['tf32']:.3f} ms, bf16 {out['bf16']:.3f} ms", flush=True)
