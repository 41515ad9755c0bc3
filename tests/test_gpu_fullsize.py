"""GPU parity at the sizes that are benchmarked (BASELINE configs C2 and C4), against the ORACLE — not
against another kernel of this library: C2 at full size on both scan families (256 random queries
each), a C4-shaped sharded search (1024-bit signatures, top-100, >= 2M rows, 3 shards, planted
neighbours and i.i.d. boundary ties), the re-rank with duplicated candidates."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu
UMAX = np.iinfo(np.uint64).max


@pytest.fixture(scope="module")
def nps():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import np_sims
    return np_sims


@pytest.mark.parametrize("family", ["mma", "popc"])
def test_c2_full_size_vs_oracle(nps, family):
    import torch
    from np_sims import _lib
    rng = np.random.default_rng(2)
    h = rng.integers(0, 2**64, size=(400_000, 10), dtype=np.uint64)
    q = rng.integers(0, 2**64, size=(10_000, 10), dtype=np.uint64)
    near = rng.integers(0, len(h), size=3000)
    q[:3000] = h[near]
    q[:3000, 0] ^= rng.integers(0, 2**16, size=3000, dtype=np.uint64)     # planted neighbours at distance <= 16
    H, Q = torch.from_numpy(h.view(np.int64)).cuda(), torch.from_numpy(q.view(np.int64)).cuda()
    sel = np.sort(rng.choice(len(q), size=256, replace=False))
    prev = _lib.set_scan_path(_lib.SCAN_MMA if family == "mma" else _lib.SCAN_POPC)
    try:
        if family == "mma":
            rows, dists = nps.hamming_top_n(H, Q, 10, return_distances=True)
            assert _lib.raw("nps_last_scan_path")() == _lib.SCAN_MMA
            rows = rows.view(torch.int64)[torch.from_numpy(sel).cuda()]
            dists = dists.view(torch.int64)[torch.from_numpy(sel).cuda()]
        else:       # the XOR+POPC family is 20x slower on this batch: only the checked queries
            rows, dists = nps.hamming_top_n(H, Q[torch.from_numpy(sel).cuda()].contiguous(), 10, return_distances=True)
            assert _lib.raw("nps_last_scan_path")() == _lib.SCAN_POPC
    finally:
        _lib.set_scan_path(prev)
    rows, dists = rows.view(torch.int64).cpu().numpy().view(np.uint64), dists.view(torch.int64).cpu().numpy().view(np.uint64)
    for j, i in enumerate(sel):
        want_r, want_d = oracle.top_k_canonical(h, q[i], 10)
        assert np.array_equal(rows[j], want_r), (family, i)
        assert np.array_equal(dists[j], want_d), (family, i)


def test_c4_shape_three_shards_vs_oracle(nps):
    """W = 16, k = 100, 2.1M rows in 3 row shards, 512 queries (tensor-core family per shard), keys
    merged by K4 exactly as the multi-GPU path does after its all-gather."""
    import torch
    from np_sims import _lib
    from np_sims.sharded import shard_bounds
    rng = np.random.default_rng(4)
    n, w, nq, k, G = 2_100_001, 16, 512, 100, 3
    h = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64)
    q = rng.integers(0, 2**64, size=(nq, w), dtype=np.uint64)
    for j in range(64):                                  # planted neighbours in every shard, a few flipped bits each
        for t in range(6):
            row = (j * 31337 + t * 700_001) % n
            h[row] = q[j]
            h[row, t % w] ^= np.uint64((1 << (t + 1)) - 1)
    H = torch.from_numpy(h.view(np.int64)).cuda()
    Q = torch.from_numpy(q.view(np.int64)).cuda()
    b = shard_bounds(n, G)
    keys = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
    for g in range(G):
        shard = H[b[g]:b[g + 1]]
        ws_bytes = _lib.raw("nps_hamming_top_n_workspace_bytes")(shard.shape[0], w, nq, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        _lib.call("nps_hamming_top_n_keys", shard.data_ptr(), shard.shape[0], w, Q.data_ptr(), nq, k, b[g],
                  keys[g].data_ptr(), ws.data_ptr(), ws_bytes, 0)
        assert _lib.raw("nps_last_scan_path")() == _lib.SCAN_MMA
    rows = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    dists = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    mws_bytes = _lib.raw("nps_top_n_merge_workspace_bytes")(nq, G, k)
    mws = torch.empty(max(mws_bytes, 16), dtype=torch.uint8, device="cuda")
    _lib.call("nps_top_n_merge", keys.data_ptr(), G, nq, k, 1, rows.data_ptr(), dists.data_ptr(), mws.data_ptr(),
              mws_bytes, 0)
    torch.cuda.synchronize()
    rows, dists = rows.cpu().numpy().view(np.uint64), dists.cpu().numpy().view(np.uint64)
    ties = 0
    for i in list(range(0, 64, 4)) + list(rng.choice(np.arange(64, nq), size=32, replace=False)):
        want_r, want_d = oracle.top_k_canonical(h, q[i], k)
        assert np.array_equal(rows[i], want_r), i
        assert np.array_equal(dists[i], want_d), i
        ties += int(want_d[-1] == want_d[-2])
    assert ties > 0                                       # i.i.d. 1024-bit rows: boundary ties are the rule


def test_rerank_duplicate_candidates_are_compacted(nps):
    rng = np.random.default_rng(6)
    h = rng.integers(0, 2**64, size=(5000, 4), dtype=np.uint64)
    q = rng.integers(0, 2**64, size=(3, 4), dtype=np.uint64)
    cand = rng.integers(0, 5000, size=(3, 60)).astype(np.uint64)
    cand[:, 10:30] = cand[:, 0:20]                        # every one of the first 20 candidates listed twice
    cand[1, 40:] = UMAX                                   # padding ids are ignored
    rows, dists = nps.hamming_rerank(h, q, cand, 25, return_distances=True)
    for i in range(3):
        uniq = np.unique(cand[i][cand[i] != UMAX])
        d = oracle.hamming(h[uniq.astype(np.int64)], q[i])
        order = np.lexsort((uniq, d))[:25]
        want_r = np.full(25, UMAX, dtype=np.uint64)
        want_d = np.full(25, UMAX, dtype=np.uint64)
        want_r[:len(order)] = uniq[order]
        want_d[:len(order)] = d[order]
        assert np.array_equal(rows[i], want_r), i
        assert np.array_equal(dists[i], want_d), i


@pytest.mark.parametrize("w", [1, 4, 5, 8])
@pytest.mark.parametrize("pair", [0, 1])
def test_narrow_signatures_large_batch_vs_oracle(nps, w, pair):
    """64..512-bit signatures get two A groups per MMA issuer, so the expanders run two tiles ahead of the
    tensor core.  At work-item boundaries that once let the query tile be overwritten under pending MMAs
    (a parity wait one phase ahead reads as "done"): a few wrong results per thousand queries, only on
    shapes with many work items per CTA — every smaller test was green.  So: 1M rows x 2000 queries,
    checked against the ORACLE, single-CTA and CTA-pair MMAs."""
    import torch
    from np_sims import _lib
    rng = np.random.default_rng(50 + w)
    n, nq, k = 1_000_000, 2000, 10
    h = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64)
    q = rng.integers(0, 2**64, size=(nq, w), dtype=np.uint64)
    H, Q = torch.from_numpy(h.view(np.int64)).cuda(), torch.from_numpy(q.view(np.int64)).cuda()
    prev = _lib.set_scan_path(_lib.SCAN_MMA)
    _lib.raw("nps_set_mma_pair")(2 if pair else 0)
    try:
        for rep in range(3):
            rows, dists = nps.hamming_top_n(H, Q, k, return_distances=True)
            assert _lib.raw("nps_last_scan_path")() == _lib.SCAN_MMA
            rows = rows.view(torch.int64).cpu().numpy().view(np.uint64)
            dists = dists.view(torch.int64).cpu().numpy().view(np.uint64)
            for i in rng.choice(nq, size=48, replace=False):
                want_r, want_d = oracle.top_k_canonical(h, q[i], k)
                assert np.array_equal(rows[i], want_r), (w, pair, rep, i)
                assert np.array_equal(dists[i], want_d), (w, pair, rep, i)
    finally:
        _lib.raw("nps_set_mma_pair")(1)
        _lib.set_scan_path(prev)


@pytest.mark.gpu
def test_hash_fixer_rings_under_pressure_vs_float64_oracle(nps):
    """K1T at a size where every CTA works through many tiles (150k x 768 -> 1024 bit): ordinary rows, whose
    unsure pairs travel through the fixer warps' shared-memory rings, interleaved with runs of degenerate rows
    (zero / tiny / huge / inf / NaN: all 1024 pairs of such a row are unsure, which overflows a ring into the global
    fix-up list while other entries are in flight).  Both routes and their mixture must give the float64 oracle's
    signatures bit for bit."""
    import np_sims.lsh as lsh
    n, dims, bits = 150_000, 768, 1024
    rng = np.random.default_rng(2024)
    X = rng.standard_normal((n, dims), dtype=np.float32)
    X[::3] *= rng.uniform(1e-3, 1e3, size=(len(X[::3]), 1)).astype(np.float32)
    bad = rng.choice(n, size=1200, replace=False)
    bad[:200] = np.arange(5000, 5200)                 # a contiguous run: whole tiles of degenerate rows
    X[bad[0:300]] = 0.0
    X[bad[300:600]] *= np.float32(1e-25)
    X[bad[600:900]] *= np.float32(1e25)
    X[bad[900:1050], 5] = np.inf
    X[bad[1050:1200], 7] = np.nan
    np.random.seed(3)
    W = lsh.create_projections(bits, dims)
    with np.errstate(all="ignore"):
        want = oracle.index_np(X, np.asarray(W))
    got = np.asarray(lsh.index(X, W))
    st = lsh.hash_stats()
    assert st["path"] == "tc-split+fixup" and not st["overflowed"]
    assert st["in_kernel"] > 0 and st["listed"] >= 200 * bits, st          # both routes were taken
    assert st["recomputed"] >= len(set(bad.tolist())) * bits, st
    diff = got ^ want
    assert not diff.any(), f"{np.count_nonzero(diff.any(axis=1))} rows differ; first {np.flatnonzero(diff.any(axis=1))[:8]}"
