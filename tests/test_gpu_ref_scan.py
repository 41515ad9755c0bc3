"""GPU parity tests of the reference-order single-query path K2R (np-sims_b200/csrc/ref_scan.cuh,
ref_scan.cu, searcher.cu): `hamming_top_10` / `hamming_top_cand` / `lsh.query` must equal the reference
queue (np_sims/ufunc.c:108-195) bit for bit — ids in slot order, history-dependent tie survivors —
on tables large enough to take the streaming scan (> 32768 rows), on every width class, with heavy
ties, with adversarial row orders that overflow the survivor segments (device-gated / host-checked
fallback), batched, and from many host threads."""
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


def rand_sigs(rng, n, w, live_bits=None):
    h = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64)
    if live_bits is not None and live_bits < 64 * w:
        mask = np.zeros(w, dtype=np.uint64)
        for b in rng.choice(64 * w, size=live_bits, replace=False):
            mask[b // 64] |= np.uint64(1) << np.uint64(b % 64)
        h &= mask
    return h


def device_call(nps, h, q, k):
    """The device-level entry (nps_hamming_top_n_reference) with tensors: (rows, scores) as NumPy."""
    import torch
    H = torch.from_numpy(h.view(np.int64)).cuda()
    Q = torch.from_numpy(np.ascontiguousarray(q).view(np.int64)).cuda()
    rows, scores = nps.hamming_top_n(H, Q, k, return_distances=True, order="reference")
    return rows.cpu().numpy().view(np.uint64), scores.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("w", [1, 2, 3, 5, 8, 10, 13, 16, 20, 24, 31, 32, 33, 40, 63, 64])
def test_streaming_reference_order_all_widths(nps, w):
    rng = np.random.default_rng(700 + w)
    n = int(rng.integers(40_000, 90_000))
    live = None if w > 2 else 40                       # narrow signatures: heavy ties
    h = rand_sigs(rng, n, w, live)
    q = rand_sigs(rng, 1, w, live)[0]
    h[rng.integers(0, n, size=20)] = q                 # exact matches scattered over the table
    for k in (10, 37, 1000):
        rows, scores = device_call(nps, h, q, k)
        want_r, want_s = oracle.top_k_queue(h, q, k)
        assert np.array_equal(rows, want_r), (w, k)
        assert np.array_equal(scores, want_s), (w, k)


@pytest.mark.parametrize("n", [32_768, 32_769, 33_000, 65_537, 300_001, 1_000_003])
def test_streaming_reference_order_row_counts(nps, n):
    rng = np.random.default_rng(n)
    h = rand_sigs(rng, n, 10, live_bits=24)            # 24 live bits: the k-th distance is a huge tie class
    q = rand_sigs(rng, 1, 10, live_bits=24)[0]
    for k in (10, 1000):
        rows, scores = device_call(nps, h, q, k)
        want_r, want_s = oracle.top_k_queue(h, q, k)
        assert np.array_equal(rows, want_r) and np.array_equal(scores, want_s), (n, k)


def test_reference_order_batched_queries(nps):
    rng = np.random.default_rng(5)
    h = rand_sigs(rng, 120_000, 10)
    q = rand_sigs(rng, 7, 10)
    q[:3] = h[[5, 60_000, 119_999]]
    rows, scores = device_call(nps, h, q, 10)
    for i in range(len(q)):
        want_r, want_s = oracle.top_k_queue(h, q[i], 10)
        assert np.array_equal(rows[i], want_r) and np.array_equal(scores[i], want_s), i


def sorted_by_distance(n, w, descending):
    """Row i at distance ~ proportional to its position: with `descending` every row beats the running
    worst, so the prefix bound is useless and the survivor segments overflow."""
    h = np.zeros((n, w), dtype=np.uint64)
    bits = 64 * w
    d = (np.arange(n, dtype=np.int64) * (bits - 1)) // n
    if descending:
        d = (bits - 1) - d
    for word in range(w):
        nb = np.clip(d - 64 * word, 0, 64).astype(np.uint64)
        full = nb >= 64
        vals = np.where(full, np.uint64(UMAX), (np.uint64(1) << np.minimum(nb, np.uint64(63))) - np.uint64(1))
        h[:, word] = vals
    return h


@pytest.mark.parametrize("descending", [True, False])
def test_adversarial_row_order_takes_the_exact_fallback(nps, descending):
    import np_sims.lsh as lsh
    from np_sims._device import attach, searcher_for
    n, w = 200_000, 4
    h = sorted_by_distance(n, w, descending)
    q = np.zeros(w, dtype=np.uint64)
    for k in (10, 1000):
        rows, scores = device_call(nps, h, q, k)           # device entry: gated K5 + dense replay
        want_r, want_s = oracle.top_k_queue(h, q, k)
        assert np.array_equal(rows, want_r) and np.array_equal(scores, want_s), (descending, k)
    hd = attach(h, None)                                   # host entry (searcher): host-checked fallback
    assert np.array_equal(nps.hamming_top_10(hd, q), oracle.top_k_queue(h, q, 10)[0])
    assert np.array_equal(nps.hamming_top_cand(hd, q), oracle.top_k_queue(h, q, 1000)[0])
    st = searcher_for(hd).stats()
    assert st["calls"] == 2
    if descending:
        assert st["fallbacks"] == 2                        # every row passes the prefix bound: overflow


def test_searcher_matches_device_path_and_oracle(nps):
    """NumPy query against a resident table (DeviceArray) = the C-level searcher (CUDA graph per slot)."""
    import np_sims.lsh as lsh
    from np_sims._device import searcher_for
    rng = np.random.default_rng(11)
    X = rng.standard_normal((50_000, 96)).astype(np.float32)
    np.random.seed(3)
    W = lsh.create_projections(256, 96)
    H = lsh.index(X, W)
    Hn = np.asarray(H)
    assert not H.flags.writeable                           # resident copy attached: in-place writes are refused
    for i in (0, 1, 777, 49_999):
        sig = oracle.index_np(X[i][None, :], np.asarray(W))[0]
        assert np.array_equal(lsh.index_one(X[i], W), sig)
        for _ in range(3):                                 # first call runs directly, later ones replay the graph
            ids, none = lsh.query(X[i], H, W)
            assert none is None
            assert np.array_equal(ids, oracle.top_k_queue(Hn, sig, 10)[0]), i
            assert np.array_equal(nps.hamming_top_10(H, sig), oracle.top_k_queue(Hn, sig, 10)[0])
            assert np.array_equal(nps.hamming_top_cand(H, sig), oracle.top_k_queue(Hn, sig, 1000)[0])
        ids64, _ = lsh.query(X[i].astype(np.float64), H, W)   # float64 vectors hash in float64 too
        assert np.array_equal(ids64, oracle.top_k_queue(Hn, oracle.index_np(X[i].astype(np.float64)[None, :], np.asarray(W))[0], 10)[0])
    st = searcher_for(H).stats()
    assert st["fallbacks"] == 0 and st["slots"] == 1
    with pytest.raises(ValueError):
        lsh.query(X[0][:50], H, W)


def test_searcher_from_many_threads(nps):
    """benchmark/glove.py:183-199: ThreadPoolExecutor, one query per task.  Every thread gets its own slot."""
    from concurrent.futures import ThreadPoolExecutor
    import np_sims.lsh as lsh
    from np_sims._device import searcher_for
    rng = np.random.default_rng(12)
    X = rng.standard_normal((100_000, 64)).astype(np.float32)
    np.random.seed(4)
    W = lsh.create_projections(640, 64)
    H = lsh.index(X, W)
    Hn = np.asarray(H)
    sigs = oracle.index_np(X[:256], np.asarray(W))
    want = [oracle.top_k_queue(Hn, sigs[i], 10)[0] for i in range(256)]
    with ThreadPoolExecutor(max_workers=16) as ex:
        for _ in range(3):
            got = list(ex.map(lambda i: lsh.query(X[i], H, W)[0], range(256)))
            for i in range(256):
                assert np.array_equal(got[i], want[i]), i
            got = list(ex.map(lambda i: nps.hamming_top_10(H, sigs[i]), range(256)))
            for i in range(256):
                assert np.array_equal(got[i], want[i]), i
    st = searcher_for(H).stats()
    assert st["calls"] == 6 * 256 and 1 <= st["slots"] <= 16 and st["fallbacks"] == 0


def test_small_tables_one_launch(nps):
    from np_sims import _lib
    rng = np.random.default_rng(13)
    for n in (0, 1, 5, 10, 11, 999, 1000, 1001, 32_767):
        h = rand_sigs(rng, n, 3, live_bits=12)
        q = rand_sigs(rng, 1, 3, live_bits=12)[0]
        for k in (10, 1000):
            before = _lib.raw("nps_launch_count")()
            rows, scores = device_call(nps, h, q, k)
            assert _lib.raw("nps_launch_count")() - before == 1
            want_r, want_s = oracle.top_k_queue(h, q, k)
            assert np.array_equal(rows, want_r) and np.array_equal(scores, want_s), (n, k)
