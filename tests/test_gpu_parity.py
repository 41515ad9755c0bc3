"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check compares the CUDA path
(through the np_sims drop-in surface or the C ABI) with the oracle / the reference's golden
vectors on the same inputs.  Integer work is compared bit for bit."""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle
from tests.golden.cases import RANDOM_TIE_CASES, random_tie_case

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


# ------------------------------------------------------------------------------ K5 distances
def test_hamming_kats(nps, kat_hamming, kat_top10):
    for i in range(int(kat_hamming["num_hamming_cases"])):
        h, q, want = kat_hamming[f"hashes_{i}"], kat_hamming[f"query_{i}"], kat_hamming[f"expected_{i}"]
        got = nps.hamming_c(h, q)
        assert got.dtype == np.uint64 and np.array_equal(got, want)
        assert np.array_equal(nps.hamming_naive(h, q), want)
    for i in range(int(kat_top10["num_cases"])):
        h, q = kat_top10[f"hashes_{i}"], kat_top10[f"query_{i}"]
        assert np.array_equal(nps.hamming_c(h, q), kat_top10[f"ref_dist_{i}"]), i


@pytest.mark.parametrize("w", [1, 2, 3, 5, 8, 10, 16, 21, 32, 33, 40, 80, 200])
def test_hamming_random_widths(nps, w):
    rng = np.random.default_rng(1000 + w)
    n = 3001
    h, q = rand_sigs(rng, n, w), rand_sigs(rng, 5, w)
    want = oracle.hamming_np(h, q)
    got = nps.hamming_c(h, q)
    assert got.shape == (5, n) and np.array_equal(got, want)
    assert np.array_equal(nps.hamming_c(h, q[2]), want[2])
    for dt in (np.uint16, np.uint32):
        from np_sims import ufuncs
        assert np.array_equal(ufuncs.hamming(h, q, out_dtype=dt), want.astype(dt))


def test_num_unshared_bits_and_bit_count(nps, kat_hamming):
    from np_sims.hamming import bit_count64
    for i in range(int(kat_hamming["num_bit_cases"])):
        bits = kat_hamming[f"bits_{i}"]
        assert np.array_equal(bit_count64(bits.copy()), np.bitwise_count(bits).astype(np.uint64))
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2**64, size=(7, 33), dtype=np.uint64)
    b = rng.integers(0, 2**64, size=33, dtype=np.uint64)
    got = nps.num_unshared_bits(a, b)
    assert got.dtype == np.uint8 and np.array_equal(got, oracle.num_unshared_bits(a, b))
    assert np.array_equal(nps.num_unshared_bits(a, np.uint64(0)), np.bitwise_count(a))


# ------------------------------------------------------------------------------ reference order
def test_top10_kats_bit_identical_to_reference(nps, kat_top10):
    # the reference's own test (test/test_hamming_top_n.py:169-173) and more: identical ARRAY
    for i in range(int(kat_top10["num_cases"])):
        h, q = kat_top10[f"hashes_{i}"], kat_top10[f"query_{i}"]
        got = nps.hamming_top_10(h, q)
        assert got.dtype == np.uint64 and got.shape == (10,)
        assert set(got.tolist()) == set(kat_top10[f"expected_{i}"].tolist()), i
        assert np.array_equal(got, kat_top10[f"ref_top10_{i}"]), i
        if f"ref_cand_{i}" in kat_top10:
            assert np.array_equal(nps.hamming_top_cand(h, q), kat_top10[f"ref_cand_{i}"]), i


def test_random_ties_bit_identical_to_reference(nps, random_ties):
    for ci, (n, w, bits_kept, seed) in enumerate(RANDOM_TIE_CASES):
        h, q = random_tie_case(n, w, bits_kept, seed)
        assert np.array_equal(nps.hamming_top_10(h, q), random_ties[f"ref_top10_{ci}"]), ci
        assert np.array_equal(nps.hamming_top_cand(h, q), random_ties[f"ref_cand_{ci}"]), ci


def test_c1_glove_top10_all_queries(nps, c1_glove):
    # BASELINE config C1: 1019 self-queries, k=10, batched through the reference-order path
    H = c1_glove["H"]
    got = nps.hamming_top_n(H, H, 10, order="reference")
    assert np.array_equal(got, c1_glove["ref_top10"])
    for j, qi in enumerate(c1_glove["q_ids"]):
        assert np.array_equal(nps.hamming_top_cand(H, H[qi]), c1_glove["ref_cand"][j])
        assert np.array_equal(nps.hamming_c(H, H[qi]), c1_glove["ref_dist"][j])
    # canonical order: exact vs the stable-argsort oracle, tie-aware vs the reference
    can, dist = nps.hamming_top_n(H, H, 10, return_distances=True)
    for i in range(len(H)):
        rows, d = oracle.top_k_canonical(H, H[i], 10)
        assert np.array_equal(can[i], rows) and np.array_equal(dist[i], d), i
        assert oracle.tie_aware_equal(oracle.hamming(H, H[i]), can[i], c1_glove["ref_top10"][i], 10)


# ------------------------------------------------------------------------------ canonical order
@pytest.mark.parametrize("w", list(range(1, 35)) + [40, 42, 80])
def test_top_n_canonical_all_widths(nps, w):
    rng = np.random.default_rng(w)
    n = int(rng.integers(2000, 9000))
    h = rand_sigs(rng, n, w, live_bits=min(64 * w, 24))       # heavy ties
    q = np.concatenate([h[rng.integers(0, n, size=3)], rand_sigs(rng, 2, w, 24)])
    for k in (1, 10, 100):
        rows, dists = nps.hamming_top_n(h, q, k, return_distances=True)
        assert rows.shape == (5, k) and rows.dtype == np.uint64
        for i in range(len(q)):
            want_r, want_d = oracle.top_k_canonical(h, q[i], k)
            assert np.array_equal(rows[i], want_r), (w, k, i)
            assert np.array_equal(dists[i], want_d), (w, k, i)


@pytest.mark.parametrize("n", [0, 1, 2, 9, 10, 11, 255, 256, 257, 1023, 1025, 4097, 70001])
@pytest.mark.parametrize("w", [1, 3, 10, 16])
def test_top_n_ragged_row_counts(nps, n, w):
    rng = np.random.default_rng(n * 31 + w)
    h = rand_sigs(rng, n, w, live_bits=20)
    q = rand_sigs(rng, 1, w, live_bits=20)[0]
    for k in (10, 1000):
        rows, dists = nps.hamming_top_n(h, q, k, return_distances=True)
        want_r, want_d = oracle.top_k_canonical(h, q, k)
        assert np.array_equal(rows, want_r) and np.array_equal(dists, want_d)
    assert np.array_equal(nps.hamming_top_10(h, q), oracle.top_k_queue(h, q, 10)[0])
    assert np.array_equal(nps.hamming_top_cand(h, q), oracle.top_k_queue(h, q, 1000)[0])


@pytest.mark.parametrize("nq", [1, 2, 3, 7, 129, 300])
def test_top_n_query_batches(nps, nq):
    rng = np.random.default_rng(nq)
    h = rand_sigs(rng, 20011, 10)
    q = rand_sigs(rng, nq, 10)
    rows = nps.hamming_top_n(h, q, 10)
    check = range(nq) if nq <= 7 else rng.choice(nq, size=12, replace=False)
    for i in check:
        assert np.array_equal(rows[i], oracle.top_k_canonical(h, q[i], 10)[0]), i


def test_top_n_adversarial_orders(nps):
    # distances strictly decreasing with the row index: every row beats the running threshold
    n, w = 6000, 2
    h = np.zeros((n, w), dtype=np.uint64)
    for i in range(n):
        bits = 127 - (i * 128) // n
        h[i, 0] = (1 << min(bits, 64)) - 1 if bits < 64 else UMAX
        h[i, 1] = (1 << (bits - 64)) - 1 if bits > 64 else 0
    q = np.zeros(w, dtype=np.uint64)
    for k in (10, 1000):
        rows, dists = nps.hamming_top_n(h, q, k, return_distances=True)
        want_r, want_d = oracle.top_k_canonical(h, q, k)
        assert np.array_equal(rows, want_r) and np.array_equal(dists, want_d)
    assert np.array_equal(nps.hamming_top_10(h, q), oracle.top_k_queue(h, q, 10)[0])
    assert np.array_equal(nps.hamming_top_cand(h, q), oracle.top_k_queue(h, q, 1000)[0])
    # all rows identical: pure index tie-break
    h[:] = 7
    assert nps.hamming_top_n(h, q, 10).tolist() == list(range(10))
    assert np.array_equal(nps.hamming_top_10(h, q), oracle.top_k_queue(h, q, 10)[0])


def test_sharded_search_equals_single_shot(nps):
    # row-sharded search + key merge (the multi-GPU data path, here on one device)
    import torch
    from np_sims import _lib
    rng = np.random.default_rng(77)
    n, w, nq, k, G = 50_003, 16, 9, 100, 4
    h = rand_sigs(rng, n, w, live_bits=40)
    q = rand_sigs(rng, nq, w, live_bits=40)
    single_r, single_d = nps.hamming_top_n(h, q, k, return_distances=True)
    H = torch.from_numpy(h.view(np.int64)).cuda()
    Q = torch.from_numpy(q.view(np.int64)).cuda()
    bounds = [n * g // G for g in range(G + 1)]
    keys = torch.empty((G, nq, k), dtype=torch.int64, device="cuda")
    for g in range(G):
        shard = H[bounds[g]:bounds[g + 1]].contiguous()
        ws_bytes = _lib.raw("nps_hamming_top_n_workspace_bytes")(shard.shape[0], w, nq, k)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        _lib.call("nps_hamming_top_n_keys", shard.data_ptr(), shard.shape[0], w, Q.data_ptr(), nq, k, bounds[g],
                  keys[g].data_ptr(), ws.data_ptr(), ws_bytes, 0)
    rows = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    dists = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    mws_bytes = _lib.raw("nps_top_n_merge_workspace_bytes")(nq, G, k)
    mws = torch.empty(max(mws_bytes, 16), dtype=torch.uint8, device="cuda")
    _lib.call("nps_top_n_merge", keys.data_ptr(), G, nq, k, 1, rows.data_ptr(), dists.data_ptr(), mws.data_ptr(),
              mws_bytes, 0)
    torch.cuda.synchronize()
    assert np.array_equal(rows.cpu().numpy().view(np.uint64), single_r)
    assert np.array_equal(dists.cpu().numpy().view(np.uint64), single_d)


# ------------------------------------------------------------------------------ K1 hashing
def test_c1_hash_bits_equal_reference(nps, c1_glove):
    import np_sims.lsh as lsh
    np.random.seed(0)
    W = lsh.create_projections(640, 300)
    X, H = c1_glove["X"], c1_glove["H"]
    got = lsh.index(X, W)
    assert got.dtype == np.uint64 and got.shape == H.shape
    flipped = int(np.bitwise_count(np.asarray(got) ^ H).sum())
    assert flipped == 0, f"{flipped} of {H.size * 64} hash bits differ from the reference"
    assert np.array_equal(lsh.index_one(X[774], W), H[774])
    # lsh.query end to end == the reference's returned arrays (np_sims/lsh.py:70-74)
    for j, qi in enumerate(c1_glove["q_ids"]):
        ids, none = lsh.query(X[qi], got, W)
        assert none is None and np.array_equal(ids, c1_glove["ref_query"][j])
        idxs, sims = lsh.query_with_hamming_then_slow_argsort(X[qi], got, W, n=10)
        assert np.array_equal(sims, c1_glove["ref_argsort_dists"][j])
        assert np.array_equal(idxs, oracle.top_k_canonical(H, H[qi], 10)[0].astype(np.int64))
        front = lsh.front_hashes(got, 128)
        two, _ = lsh.query_two_phase(X[qi], front, got, W)
        assert np.array_equal(two, c1_glove["ref_two_phase"][j])


@pytest.mark.parametrize("dims,bits,dtype", [(50, 64, np.float64), (300, 640, np.float32), (768, 1024, np.float32),
                                             (13, 128, np.float32)])
def test_hash_random_vs_oracle(nps, dims, bits, dtype):
    import np_sims.lsh as lsh
    rng = np.random.default_rng(dims)
    X = rng.standard_normal((777, dims)).astype(dtype)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    np.random.seed(3)
    W = lsh.create_projections(bits, dims)
    want = oracle.index(X, np.asarray(W))
    got = np.asarray(lsh.index(X, W))
    diff = np.bitwise_count(got ^ want).sum()
    if diff:   # only summation-order noise is allowed: |x.w| at the 1e-15 level
        dots = oracle.project(X, np.asarray(W))
        bad = np.unpackbits((got ^ want).view(np.uint8), axis=1, bitorder="little").astype(bool)
        assert np.abs(dots[bad]).max() < 1e-14
    with pytest.raises(ValueError, match="multiple of 64"):
        lsh.index(X, np.asarray(W)[:-1])


@pytest.mark.parametrize("seed", [0, 1000])
@pytest.mark.parametrize("num_projections", [64, 128, 256, 2560])
@pytest.mark.parametrize("num_dims", [50, 100, 300])
@pytest.mark.parametrize("num_fronts", [None, 64, 256])
def test_index_and_query_gets_more_similar(nps, seed, num_projections, num_dims, num_fronts):
    # mirrors reference test/test_lsh.py:45-73 (including its skip condition)
    import np_sims.lsh as lsh
    if num_fronts is not None and num_fronts <= num_projections:
        pytest.skip("Skipping illegal state (we should test another place")
    np.random.seed(seed)
    projections = lsh.create_projections(num_projections, dims=num_dims)
    query = np.zeros(num_dims)
    query[0], query[1] = 0.8, 0.2
    query /= np.linalg.norm(query)
    all_vectors, similar, different = [], [], []
    for _ in range(100):
        alike, diff = np.zeros(num_dims), np.zeros(num_dims)
        alike[0], diff[-1] = 1.0, 1.0
        similar.append(len(all_vectors)); all_vectors.append(alike)
        different.append(len(all_vectors)); all_vectors.append(diff)
    hashed = lsh.index(np.array(all_vectors), projections)
    if num_fronts is None:
        results, _ = lsh.query(query, hashed, projections)
    else:
        results, _ = lsh.query_two_phase(query, lsh.front_hashes(hashed, num_fronts), hashed, projections)
    for r in results:
        assert r in similar and r not in different
    # and the whole thing equals the oracle's restatement of the same pipeline
    H = oracle.index(np.array(all_vectors), np.asarray(projections))
    assert np.array_equal(np.asarray(hashed), H)


# ------------------------------------------------------------------------------ C ABI, host level
def test_c_abi_host_entry_points(nps, kat_top10):
    from np_sims import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.nps_last_error.restype = ctypes.c_char_p
    vp = ctypes.c_void_p
    for i in (0, 7, 13, 14, 20, 22):
        h = np.ascontiguousarray(kat_top10[f"hashes_{i}"])
        q = np.ascontiguousarray(kat_top10[f"query_{i}"])
        out = np.zeros(10, dtype=np.uint64)
        rc = lib.nps_host_hamming_top_k(vp(h.ctypes.data), vp(q.ctypes.data), ctypes.c_uint64(h.shape[0]),
                                        ctypes.c_uint64(h.shape[1]), ctypes.c_uint32(10), vp(out.ctypes.data))
        assert rc == 0, lib.nps_last_error()
        assert np.array_equal(out, kat_top10[f"ref_top10_{i}"]), i
        d = np.zeros(h.shape[0], dtype=np.uint64)
        rc = lib.nps_host_hamming(vp(h.ctypes.data), vp(q.ctypes.data), ctypes.c_uint64(h.shape[0]),
                                  ctypes.c_uint64(h.shape[1]), vp(d.ctypes.data))
        assert rc == 0 and np.array_equal(d, kat_top10[f"ref_dist_{i}"])
    # resident index, batched queries, both orders
    rng = np.random.default_rng(9)
    h = rand_sigs(rng, 30_001, 10, live_bits=30)
    q = rand_sigs(rng, 33, 10, live_bits=30)
    ix = vp()
    assert lib.nps_index_create(vp(h.ctypes.data), ctypes.c_uint64(len(h)), 10, 0, ctypes.byref(ix)) == 0
    rows = np.zeros((33, 10), dtype=np.uint64)
    dists = np.zeros((33, 10), dtype=np.uint64)
    assert lib.nps_index_query_top_n(ix, vp(q.ctypes.data), 33, 10, 0, vp(rows.ctypes.data), vp(dists.ctypes.data)) == 0
    for i in range(33):
        wr, wd = oracle.top_k_canonical(h, q[i], 10)
        assert np.array_equal(rows[i], wr) and np.array_equal(dists[i], wd)
    assert lib.nps_index_query_top_n(ix, vp(q.ctypes.data), 33, 10, 1, vp(rows.ctypes.data), None) == 0
    for i in range(33):
        assert np.array_equal(rows[i], oracle.top_k_queue(h, q[i], 10)[0])
    assert lib.nps_index_query_top_n(ix, vp(q.ctypes.data), 33, 5000, 0, vp(rows.ctypes.data), None) == -1
    assert b"k=5000" in lib.nps_last_error()
    lib.nps_index_destroy(ix)


def test_error_behaviour_matches_reference(nps):
    h = np.zeros((4, 2), dtype=np.uint64)
    with pytest.raises(TypeError):                       # NumPy: no loop for int64 (ufunc.c:564)
        nps.hamming_top_10(h.astype(np.int64), h[0])
    with pytest.raises(ValueError, match="core dimension"):
        nps.hamming_top_10(h, np.zeros(3, dtype=np.uint64))
    with pytest.raises(ValueError):
        nps.hamming_top_n(h, h[0], 0)
    # non-contiguous input is honoured (the reference silently mis-reads strides)
    rng = np.random.default_rng(3)
    big = rand_sigs(rng, 500, 6)
    view = big[:, ::2]
    assert np.array_equal(nps.hamming_top_n(view, view[5], 10), oracle.top_k_canonical(view, view[5], 10)[0])


def test_torch_tensor_inputs_stay_on_device(nps):
    import torch
    rng = np.random.default_rng(4)
    h, q = rand_sigs(rng, 5000, 10), rand_sigs(rng, 4, 10)
    H = torch.from_numpy(h.view(np.int64)).cuda()
    Q = torch.from_numpy(q.view(np.int64)).cuda()
    rows = nps.hamming_top_n(H, Q, 10)
    assert rows.is_cuda and rows.dtype == torch.uint64
    got = rows.view(torch.int64).cpu().numpy().view(np.uint64)
    for i in range(4):
        assert np.array_equal(got[i], oracle.top_k_canonical(h, q[i], 10)[0])


# ------------------------------------------------------------------------------ K1T tensor-core hashing
@pytest.mark.parametrize("n,dims,bits", [(1, 300, 640), (127, 300, 640), (1019, 300, 640), (5000, 768, 1024),
                                         (3001, 100, 128), (2000, 52, 64), (4096, 300, 2560),
                                         (200, 8, 64), (333, 36, 128)])       # TMA boxes wider than the matrix
def test_tensor_core_hash_equals_float64_oracle(nps, n, dims, bits):
    """K1T (tcgen05 GEMM on two-term splits of both operands + float64 fix-up) must give the oracle's
    float64 signatures with either operand encoding — BF16 x 2 (default) and TF32 x 2; the bare
    tensor-core signs may only differ where |x.w| is below
    (c + dims / 4) * 2^-22 * |x| * max|w|, c = 52 (BF16) / 8 (TF32) — the TF32 kernel's tolerance and an upper
    bound of the BF16 kernel's (which uses prefix norms for the accumulation term, project_tc.cu)."""
    import torch
    import np_sims.lsh as lsh
    from np_sims import _lib
    rng = np.random.default_rng(n + dims)
    X = rng.standard_normal((n, dims)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    X[: n // 7] *= rng.uniform(0.01, 50.0, size=(n // 7, 1)).astype(np.float32)      # not all unit norm
    np.random.seed(bits)
    W = lsh.create_projections(bits, dims)
    want = oracle.index_np(X, np.asarray(W))
    dots = X.astype(np.float64) @ np.asarray(W).T
    wmax = np.linalg.norm(np.asarray(W), axis=1).max()
    for path, name, c in ((_lib.HASH_BF16, "bf16x2", 52), (_lib.HASH_TF32, "tf32x2", 8)):
        _lib.call("nps_set_hash_path", path)
        try:
            got = lsh.index(X, W)
            st = lsh.hash_stats()
            assert st["path"] == "tc-split+fixup" and not st["overflowed"]
            assert np.array_equal(np.asarray(got), want), name
            frac = st["recomputed"] / st["pairs"]
            assert frac < 0.02, frac
            # bare tensor-core pass: flipped bits only where the float64 projection is inside the tolerance
            raw = lsh.hash_vectors(torch.from_numpy(X).cuda(), W, tf32_raw=True).cpu().numpy().view(np.uint64)
        finally:
            _lib.call("nps_set_hash_path", _lib.HASH_AUTO)
        flips = np.unpackbits((raw ^ want).view(np.uint8), bitorder="little").reshape(n, -1)[:, :bits].astype(bool)
        tol = (c + dims / 4) * 2.0 ** -22 * np.linalg.norm(X.astype(np.float64), axis=1, keepdims=True) * wmax * 1.0001
        assert not np.any(flips & (np.abs(dots) >= tol)), f"{name}: flipped a bit outside the stated tolerance"
        used = float((np.abs(dots) / tol)[flips].max()) if flips.any() else 0.0
        print(f"K1T {name} n={n} dims={dims} bits={bits}: raw flipped-bit rate {flips.mean():.2e} (largest flipped "
              f"|x.w| = {used:.3f} of the tolerance), recomputed in float64 {frac:.3%}")


@pytest.mark.gpu
@pytest.mark.parametrize("structure", ["iid", "decaying", "front_loaded", "sparse_w"])
def test_bf16_hash_tolerance_is_the_documented_one(nps, structure):
    """The BF16 kernel trusts an accumulator when |C1 + C2| >= (52 + 0.006 D) 2^-22 |x| max|w| + 2.625 * 1.02 * 2^-22 * P,
    P = sum_i |x[0 : 16 (i+1))| max_j |w_j[0 : 16 (i+1))| (project_tc.cu).  Actual tensor-core errors are ~1000 x smaller
    than that, so a kernel whose tolerance came out too SMALL would still hash correctly in every other test: here
    the number of pairs it recomputes in float64 is compared with the number the documented formula selects (NumPy,
    float64), on rows and projections with and without structure along the dimension axis."""
    import np_sims.lsh as lsh
    n, dims, bits = 4096, 768, 1024
    rng = np.random.default_rng(11)
    X = rng.standard_normal((n, dims)).astype(np.float32)
    np.random.seed(5)
    W = np.asarray(lsh.create_projections(bits, dims)).copy()
    if structure == "decaying":
        X *= np.exp(-np.arange(dims) / 100.0).astype(np.float32)
    elif structure == "front_loaded":
        X[:, :16] *= 50.0
        X[:, 400:] = 0.0
    elif structure == "sparse_w":
        W[:, rng.permutation(dims)[: dims // 2]] = 0.0
        W *= rng.uniform(0.5, 2.0, size=(bits, 1))
    want = oracle.index_np(X, W)
    got = np.asarray(lsh.index(X, W))
    st = lsh.hash_stats()
    assert st["path"] == "tc-split+fixup" and not st["overflowed"]
    assert np.array_equal(got, want)
    x64, nblk = X.astype(np.float64), dims // 16
    dots = x64 @ W.T
    px = np.sqrt(np.cumsum((x64 ** 2).reshape(n, nblk, 16).sum(axis=2), axis=1))                  # [n, nblk]
    qtab = np.sqrt(np.cumsum((W ** 2).reshape(bits, nblk, 16).sum(axis=2), axis=1)).max(axis=0)   # [nblk]
    wmax = np.linalg.norm(W, axis=1).max()
    tol = ((52 + 0.006 * dims) * 2.0 ** -22 * np.linalg.norm(x64, axis=1) * wmax + 2.625 * 1.02 * 2.0 ** -22 * (px @ qtab))
    lo = np.count_nonzero(np.abs(dots) < tol[:, None] * 0.995)
    hi = np.count_nonzero(np.abs(dots) < tol[:, None] * 1.005)
    assert lo <= st["recomputed"] <= hi, (structure, lo, st["recomputed"], hi)
    full = np.count_nonzero(np.abs(dots) < ((52 + dims / 4) * 2.0 ** -22 * np.linalg.norm(x64, axis=1) * wmax)[:, None])
    print(f"K1T bf16 {structure}: {st['recomputed']} pairs recomputed ({st['recomputed'] / dots.size:.3%}); "
          f"the full-norm (52 + D/4) tolerance of round 1 would have selected {full}")


def test_tensor_core_hash_degenerate_inputs(nps):
    import np_sims.lsh as lsh
    np.random.seed(1)
    W = lsh.create_projections(128, 64)
    X = np.zeros((300, 64), dtype=np.float32)                   # zero vectors: dot == 0 -> every bit 1
    X[100:200] = np.asarray(W)[5].astype(np.float32) * 0.0      # still zero
    X[200:] = np.eye(64, dtype=np.float32)[np.arange(100) % 64] * 1e-20
    got = lsh.index(X, W)
    assert np.array_equal(np.asarray(got), oracle.index_np(X, np.asarray(W)))
    st = lsh.hash_stats()           # every row is zero or tiny: all pairs listed, list overflows, and the
    assert st["path"] == "tc-split+fixup" and st["overflowed"]      # gated float64 launch redoes the call
    # a few wild rows among ordinary ones: only those rows go to the fix-up list
    rng = np.random.default_rng(9)
    X3 = rng.standard_normal((5000, 64)).astype(np.float32)
    X3[7] = 0.0
    X3[11] *= 1e-25
    X3[13] *= 1e25
    X3[17, 3] = np.inf
    X3[19, 5] = np.nan
    with np.errstate(all="ignore"):
        want3 = oracle.index_np(X3, np.asarray(W))
    assert np.array_equal(np.asarray(lsh.index(X3, W)), want3)
    st = lsh.hash_stats()
    assert not st["overflowed"] and st["recomputed"] >= 5 * 128
    # float64 vectors and dims % 4 != 0 take the float64 kernel
    X2 = np.random.default_rng(0).standard_normal((50, 63))
    W2 = lsh.create_projections(64, 63)
    assert np.array_equal(np.asarray(lsh.index(X2, W2)), oracle.index_np(X2, np.asarray(W2)))
    assert lsh.last_hash_stats["path"] == "f64"


@pytest.mark.gpu
def test_replicated_index_query_slices_equal_single_shot(nps):
    # query-sharded mode (table replicated): answering the batch slice by slice == one call
    import torch
    from np_sims.sharded import ReplicatedIndex, query_bounds
    rng = np.random.default_rng(78)
    n, w, nq, k, G = 20_011, 10, 301, 10, 3
    h = rand_sigs(rng, n, w, live_bits=40)
    q = rand_sigs(rng, nq, w, live_bits=40)
    want_r, want_d = nps.hamming_top_n(h, q, k, return_distances=True)
    ix = ReplicatedIndex(torch.from_numpy(h.view(np.int64)).cuda())
    Q = torch.from_numpy(q.view(np.int64)).cuda()
    b = query_bounds(nq, G)
    parts = [ix.query_local(Q[b[g]:b[g + 1]], k) for g in range(G)]
    rows = torch.cat([p[0].view(torch.int64) for p in parts]).cpu().numpy().view(np.uint64)
    dists = torch.cat([p[1].view(torch.int64) for p in parts]).cpu().numpy().view(np.uint64)
    assert np.array_equal(rows, want_r) and np.array_equal(dists, want_d)
    full_r, full_d = ix.query(Q, k)          # world size 1: gather is the identity
    assert np.array_equal(full_r.cpu().numpy().view(np.uint64), want_r)


@pytest.mark.gpu
def test_merge_more_than_65535_queries(nps):
    # the C4 batch is 100k queries: the merge grid must not put queries on gridDim.y (max 65535)
    import torch
    from np_sims.sharded import _cuda_merge
    rng = np.random.default_rng(79)
    G, q, k = 3, 70_001, 4
    keys = np.sort(rng.integers(0, 2**50, size=(G, q, k), dtype=np.uint64), axis=2)
    rows, dists = _cuda_merge(torch.from_numpy(keys.view(np.int64)).cuda(), k)
    want = np.sort(np.transpose(keys, (1, 0, 2)).reshape(q, -1), axis=1)[:, :k]
    assert np.array_equal(rows.cpu().numpy().view(np.uint64), want & np.uint64((1 << 40) - 1))
    assert np.array_equal(dists.cpu().numpy().view(np.uint64), want >> np.uint64(40))


# ------------------------------------------------------------------------------ K6 two-phase, batched
def test_rerank_and_batched_two_phase_equal_oracle(nps, c1_glove):
    """lsh.py:84-91 as a batch: front-word top-1000 (canonical) -> K6 re-rank in place == the
    NumPy restatement, on the reference's glove sample and on random signatures with padding."""
    import torch
    import np_sims.lsh as lsh
    H = c1_glove["H"]
    q_ids = c1_glove["q_ids"]
    front = np.ascontiguousarray(H[:, :2])
    cand = nps.hamming_top_n(front, np.ascontiguousarray(H[q_ids][:, :2]), 1000)
    rows, dists = nps.hamming_rerank(H, H[q_ids], cand, 10, return_distances=True)
    for j, qi in enumerate(q_ids):
        want_r, want_d = oracle.two_phase_canonical(H, H[qi], 2, 1000, 10)
        assert np.array_equal(rows[j], want_r) and np.array_equal(dists[j], want_d)
        full = oracle.hamming(H, H[qi])                    # same distances as the reference's two-phase
        assert np.array_equal(np.sort(full[c1_glove["ref_two_phase"][j].astype(np.int64)]), dists[j])
    # vectors in, ids out (hash -> front scan -> re-rank), batch of 200 queries, tensor-core scan
    X = c1_glove["X"]
    np.random.seed(0)
    W = lsh.create_projections(640, 300)
    got = lsh.query_two_phase_batch(X[:200], front, H, W, n=10, num_candidates=1000)
    sig = oracle.index_np(X[:200], np.asarray(W))
    for i in (0, 7, 199):
        assert np.array_equal(got[i], oracle.two_phase_canonical(H, sig[i], 2, 1000, 10)[0])
    # fewer rows than candidates (padding ids), odd width, k > rows
    rng = np.random.default_rng(81)
    h = rand_sigs(rng, 37, 3)
    q = rand_sigs(rng, 5, 3)
    cand = nps.hamming_top_n(np.ascontiguousarray(h[:, :1]), np.ascontiguousarray(q[:, :1]), 64)
    assert cand[0, 37] == np.iinfo(np.uint64).max
    rows, dists = nps.hamming_rerank(h, q, cand, 40, return_distances=True)
    for i in range(5):
        want_r, want_d = oracle.two_phase_canonical(h, q[i], 1, 64, 40)
        assert np.array_equal(rows[i], want_r) and np.array_equal(dists[i], want_d)
    # device tensors stay on the device
    out = nps.hamming_rerank(torch.from_numpy(h.view(np.int64)).cuda(), torch.from_numpy(q.view(np.int64)).cuda(),
                             torch.from_numpy(cand.view(np.int64)).cuda(), 10)
    assert out.is_cuda and np.array_equal(out.cpu().numpy().view(np.uint64), rows[:, :10])


def test_recall_harness_index_files_use_the_reference_format(nps, tmp_path, monkeypatch):
    """benchmark/recall.py caches the index under the reference's file names and formats
    (benchmark/glove.py:262-279): np.save of uint64 [N, B/64] and float64 [B, D]."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "recall_harness", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmark", "recall.py"))
    rec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rec)
    monkeypatch.setattr(rec, "DATA_PATH", str(tmp_path) + "/")
    x = rec.synthetic_vectors(700, 64, 20, 3)
    q, ix = rec.test_train_split(x, 50, np.random.default_rng(0))
    assert q.shape == (50, 64) and ix.shape == (650, 64)
    hashes, projs, _ = rec.load_or_build_index(ix, 128, 0)
    hf, pf = tmp_path / "hashes_650_128_64_0.npy", tmp_path / "projs_650_128_64_0.npy"
    assert hf.exists() and pf.exists()
    h2, p2, _ = rec.load_or_build_index(ix, 128, 0)                     # second call loads the files
    assert h2.dtype == np.uint64 and h2.shape == (650, 2) and p2.dtype == np.float64 and p2.shape == (128, 64)
    assert np.array_equal(h2, np.asarray(hashes)) and np.array_equal(h2, oracle.index_np(ix, p2))
    gt = rec.most_similar_cos(ix, q)
    ids, _ = rec.run_algorithm("lsh_batch", q, h2, p2, None)
    assert ids.shape == (50, 10) and 0.0 <= rec.recall_at_10(ids, gt) <= 1.0
    one, _ = rec.run_algorithm("lsh_pure_c", q[:3], h2, p2, None)
    full = oracle.hamming(h2, oracle.index_np(q[:1], p2)[0])
    assert oracle.tie_aware_equal(full, one[0], ids[0], 10)


def test_query_stream_pipeline_equals_query_batch(nps, c1_glove):
    """np_sims.pipeline.QueryStream: batches streamed with copies overlapped == lsh.query_batch per batch."""
    import np_sims.lsh as lsh
    from np_sims.pipeline import QueryStream
    X, H = c1_glove["X"], c1_glove["H"]
    np.random.seed(0)
    W = lsh.create_projections(640, 300)
    batches = [X[i * 100:(i + 1) * 100 + (7 if i == 3 else 0)] for i in range(6)]      # one ragged batch
    qs = QueryStream(H, W, n=10, depth=2)
    got = [ids.copy() for ids in qs.map(batches)]
    assert len(got) == len(batches)
    for b, ids in zip(batches, got):
        assert np.array_equal(ids, lsh.query_batch(b, H, W, n=10))
    t0 = qs.submit(batches[0])
    assert np.array_equal(qs.result(t0), got[0])
    with pytest.raises(ValueError):
        qs.result(t0 - 5)


def test_single_query_api_from_a_thread_pool(nps, c1_glove):
    """The reference drives its single-query API from a ThreadPoolExecutor (benchmark/glove.py:183-199,
    GIL released in the ufunc).  The drop-in must give the same answers when called that way."""
    from concurrent.futures import ThreadPoolExecutor
    import np_sims.lsh as lsh
    X, H = c1_glove["X"], c1_glove["H"]
    np.random.seed(0)
    W = lsh.create_projections(640, 300)
    Hd = lsh.index(X, W)                                   # DeviceArray: table stays resident
    assert np.array_equal(np.asarray(Hd), H)
    want = [np.asarray(lsh.query(X[i], Hd, W)[0]) for i in range(64)]
    with ThreadPoolExecutor(max_workers=8) as ex:
        got = list(ex.map(lambda i: np.asarray(lsh.query(X[i], Hd, W)[0]), range(64)))
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
    with ThreadPoolExecutor(max_workers=8) as ex:
        got = list(ex.map(lambda i: nps.hamming_top_n(Hd, H[i], 10), range(64)))
    for i, b in enumerate(got):
        assert np.array_equal(b, oracle.top_k_canonical(H, H[i], 10)[0])
