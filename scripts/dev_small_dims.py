"""Development aid (GPU box): K1T with very small dims (TMA boxes wider than the tensor)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims.lsh as lsh
from oracle import oracle
for d in (4, 8, 12, 16, 28, 32, 36, 60, 64, 68):
    rng = np.random.default_rng(d)
    X = rng.standard_normal((333, d)).astype(np.float32)
    np.random.seed(d)
    W = lsh.create_projections(128, d)
    try:
        got = np.asarray(lsh.index(X, W))
        st = lsh.hash_stats()
        print(d, st["path"], "equal", bool(np.array_equal(got, oracle.index_np(X, np.asarray(W)))), st.get("recomputed"), flush=True)
    except Exception as e:
        print(d, "ERROR", str(e)[:150], flush=True)
