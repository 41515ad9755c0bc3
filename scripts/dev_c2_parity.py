"""Development aid (GPU box): the bench's C2 step, tensor-core vs POPC scan vs oracle hashes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
from np_sims import _lib
import bench
from oracle import oracle
W = bench.make_projections(640, 300)
X = bench.make_vectors(400_000, 300, 0)
Qv = bench.make_vectors(10_000, 300, 1)
H = lsh.hash_vectors(torch.from_numpy(X).cuda(), W)
Hx = lsh.hash_vectors(torch.from_numpy(X).cuda(), W, exact_only=True)
print("index K1T == f64:", bool(torch.equal(H, Hx)), "differing words", int((H != Hx).sum()))
S = lsh.hash_vectors(torch.from_numpy(Qv).cuda(), W)
Sx = lsh.hash_vectors(torch.from_numpy(Qv).cuda(), W, exact_only=True)
print("query K1T == f64:", bool(torch.equal(S, Sx)), "differing words", int((S != Sx).sum()))
so = oracle.index_np(Qv[:64], np.asarray(W))
print("query f64 kernel == oracle (64):", bool(np.array_equal(Sx[:64].cpu().numpy().view(np.uint64), so)))
for trial in range(3):
    _lib.set_scan_path(_lib.SCAN_MMA); r1 = np_sims.hamming_top_n(Hx, Sx, 10); torch.cuda.synchronize()
    _lib.set_scan_path(_lib.SCAN_POPC); r0 = np_sims.hamming_top_n(Hx, Sx, 10); torch.cuda.synchronize()
    _lib.set_scan_path(_lib.SCAN_AUTO)
    bad = (r0.view(torch.int64) != r1.view(torch.int64)).any(dim=1)
    print("trial", trial, "queries where mma != popc:", int(bad.sum()), bad.nonzero()[:8].flatten().tolist(), flush=True)
