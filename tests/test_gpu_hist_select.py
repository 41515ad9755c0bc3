"""GPU parity tests of K2H (np-sims_b200/csrc/hist_select.cuh, hist_select.cu): canonical top-k for
few queries and large k by counting on the distance histogram.  Must equal the oracle's canonical
order (ascending (distance, row id); np_sims/ufunc.c:174-195 keeps the k smallest distances) bit for
bit: every width class, heavy ties (the k-th distance a tie class of thousands of rows), fewer than k
rows, identical rows, shard offsets, query tiles."""
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
    from np_sims import _lib
    prev = _lib.set_scan_path(_lib.SCAN_POPC)      # few-query family (the tensor-core phases are tested elsewhere)
    yield np_sims
    _lib.set_scan_path(prev)


def rand_sigs(rng, n, w, live_bits=None):
    h = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64)
    if live_bits is not None and live_bits < 64 * w:
        mask = np.zeros(w, dtype=np.uint64)
        for b in rng.choice(64 * w, size=live_bits, replace=False):
            mask[b // 64] |= np.uint64(1) << np.uint64(b % 64)
        h &= mask
    return h


def check(nps, h, q, k, queries=None, row_base=0):
    rows, dists = nps.hamming_top_n(h, q, k, return_distances=True, row_base=row_base)
    q2 = q if q.ndim == 2 else q[None, :]
    rows2 = rows if q.ndim == 2 else rows[None, :]
    dists2 = dists if q.ndim == 2 else dists[None, :]
    for i in (range(len(q2)) if queries is None else queries):
        want_r, want_d = oracle.top_k_canonical(h, q2[i], k)
        if row_base:
            want_r = np.where(want_r == UMAX, UMAX, want_r + np.uint64(row_base))
        assert np.array_equal(rows2[i], want_r), (k, i, rows2[i][:8], want_r[:8])
        assert np.array_equal(dists2[i], want_d), (k, i)


@pytest.mark.parametrize("w", [1, 2, 3, 4, 7, 8, 10, 13, 16, 21, 32, 33, 40, 63])
def test_hist_select_all_widths(nps, w):
    rng = np.random.default_rng(900 + w)
    n = int(rng.integers(20_000, 70_000))
    h = rand_sigs(rng, n, w)
    q = rand_sigs(rng, 3, w)
    q[0] = h[n // 2]
    for k in (33, 100, 1000, 2048):
        check(nps, h, q, k)


@pytest.mark.parametrize("n", [4096, 4097, 32_767, 32_768, 32_769, 100_003, 1_000_001])
def test_hist_select_row_counts_and_ties(nps, n):
    rng = np.random.default_rng(n)
    h = rand_sigs(rng, n, 10, live_bits=16)         # 17 distinct distances: the k-th one is a class of thousands
    q = rand_sigs(rng, 2, 10, live_bits=16)
    for k in (64, 1000):
        check(nps, h, q, k)


def test_hist_select_fewer_rows_than_k_and_identical_rows(nps):
    rng = np.random.default_rng(1)
    h = rand_sigs(rng, 4100, 5)
    q = rand_sigs(rng, 1, 5)[0]
    rows, dists = nps.hamming_top_n(h[:4096], q, 2048, return_distances=True)
    want_r, want_d = oracle.top_k_canonical(h[:4096], q, 2048)
    assert np.array_equal(rows, want_r) and np.array_equal(dists, want_d)
    same = np.tile(h[:1], (50_000, 1))
    rows = nps.hamming_top_n(same, q, 1000)
    assert rows.tolist() == list(range(1000))         # pure index tie-break
    few = h[:4100]
    rows, dists = nps.hamming_top_n(few, q, 2048, return_distances=True)
    want_r, want_d = oracle.top_k_canonical(few, q, 2048)
    assert np.array_equal(rows, want_r) and np.array_equal(dists, want_d)


@pytest.mark.parametrize("nq", [1, 2, 5, 8, 9, 16, 32])
def test_hist_select_query_tiles_and_shard_offset(nps, nq):
    rng = np.random.default_rng(40 + nq)
    h = rand_sigs(rng, 150_001, 16)
    q = rand_sigs(rng, nq, 16)
    q[0] = h[149_999]
    check(nps, h, q, 100, row_base=7_000_000_000)
    check(nps, h, q, 1000, queries=[0, nq - 1])


def test_hist_select_is_the_path_taken(nps):
    from np_sims import _lib
    import torch
    rng = np.random.default_rng(3)
    H = torch.from_numpy(rand_sigs(rng, 50_000, 10).view(np.int64)).cuda()
    Q = torch.from_numpy(rand_sigs(rng, 1, 10).view(np.int64)).cuda()
    before = _lib.raw("nps_launch_count")()
    nps.hamming_top_n(H, Q, 1000)
    assert _lib.raw("nps_launch_count")() - before == 3      # H1 scan, H2 count, H3 emit + sort
    before = _lib.raw("nps_launch_count")()
    nps.hamming_top_n(H, Q, 10)
    assert _lib.raw("nps_launch_count")() - before == 3      # small k takes the same path
    small = H[:1000].contiguous()
    before = _lib.raw("nps_launch_count")()
    nps.hamming_top_n(small, Q, 10)
    assert _lib.raw("nps_launch_count")() - before == 2      # tiny table: K2 scan + K3 merge
