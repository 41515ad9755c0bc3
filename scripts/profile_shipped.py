"""One short pass over every kernel that ships, at the sizes the bench quotes, for ncu:

    python scripts/profile_shipped.py            # plain run (must exit 0 before ncu is used)
    ncu --set full --clock-control none --import-source on -k "regex:hamming_dist|hist_|merge_|mma_|project_|ref_|replay_|rerank|scan_topk" -c 70 -o gpurun_out/r02_shipped \
        python scripts/profile_shipped.py

Launch order (each kernel family once, no warm-up: ncu replays every launch anyway):
  1. K1T   project_bf16_kernel<128> + project_fix_kernel + gated float64: 2M x 768 fp32 -> 1024 bit (C3 shape, 1/5 of the rows)
  2. K2T   mma_scan_kernel (7 phases) + mma_select_warp_kernel: C2 (400k x 640 bit, 10k queries, top-10)
  3. K2T   mma_scan_kernel, CTA pairs + mma_select_kernel: C4 shape on 2M rows (1024 bit, 4096 queries, top-100)
  4. K2H   hist_scan_kernel<16,2> + hist_count_kernel + hist_emit_kernel: 1 query x 25M x 1024 bit, top-10
  5. K2R   ref_prefix_kernel + ref_scan_kernel<16,2> (hamming_top_10, reference order) on the same table
  6. K2R   ref_prefix_kernel + ref_scan_kernel<10,2> + project_one_kernel through the searcher (lsh.query), C2 table
  7. K5    hamming_dist_kernel, K6 rerank_kernel, K4 merge (row-sharded keys), K2 + K3 on a table too small for K2H
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
import np_sims.lsh as lsh
from np_sims import _lib
from np_sims.sharded import ShardedIndex

g = torch.Generator(device="cuda"); g.manual_seed(1)
# 1. K1T
X = torch.empty((2_000_000, 768), dtype=torch.float32, device="cuda").normal_(generator=g)
np.random.seed(0)
W3 = lsh.create_projections(1024, 768)
lsh.hash_vectors(X, W3)
del X
# 2. K2T single-CTA, C2
H = torch.randint(-2**63, 2**63 - 1, (400_000, 10), dtype=torch.int64, device="cuda", generator=g)
Q = torch.randint(-2**63, 2**63 - 1, (10_000, 10), dtype=torch.int64, device="cuda", generator=g)
np_sims.hamming_top_n(H, Q, 10)
# 3. K2T CTA pairs, C4 shape
H4 = torch.randint(-2**63, 2**63 - 1, (2_000_000, 16), dtype=torch.int64, device="cuda", generator=g)
Q4 = torch.randint(-2**63, 2**63 - 1, (4096, 16), dtype=torch.int64, device="cuda", generator=g)
np_sims.hamming_top_n(H4, Q4, 100)
assert _lib.last_mma_plan()["pair"] & 1
# 7a. K4 merge of two shards' keys
a, b = ShardedIndex(H4[:1_000_000], 0), ShardedIndex(H4[1_000_000:], 1_000_000)
keys = torch.stack([a._keys_fn(a.hashes, Q4[:512].contiguous(), 100, 0), b._keys_fn(b.hashes, Q4[:512].contiguous(), 100, 1_000_000)])
a._merge_fn(keys, 100)
del H4, Q4
# 4. K2H stream
Hb = torch.randint(-2**63, 2**63 - 1, (25_000_000, 16), dtype=torch.int64, device="cuda", generator=g)
qb = torch.randint(-2**63, 2**63 - 1, (1, 16), dtype=torch.int64, device="cuda", generator=g)
np_sims.hamming_top_n(Hb, qb, 10)
# 5. K2R on the same table
np_sims.hamming_top_n(Hb, qb[0], 10, order="reference")
del Hb
# 6. searcher: lsh.query over the C2 table
rng = np.random.default_rng(0)
Xc = rng.standard_normal((400_000, 300), dtype=np.float32)
np.random.seed(0)
W2 = lsh.create_projections(640, 300)
Hc = lsh.index(Xc, W2)
for i in range(3):
    lsh.query(Xc[i], Hc, W2)
# 7b. K5, K6, K2 + K3
np_sims.hamming_c(H, Q[0])
cand = np_sims.hamming_top_n(H[:, :2].contiguous(), Q[:64, :2].contiguous(), 100)
np_sims.hamming_rerank(H, Q[:64].contiguous(), cand, 10)
np_sims.hamming_top_n(H[:3000].contiguous(), Q[:4].contiguous(), 10)
torch.cuda.synchronize()
print("profile workload done")
