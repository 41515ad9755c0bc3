"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE in this container.

    python tests/golden/make_golden.py        (needs /root/reference and `make -C oracle ref`)

What is executed is the reference's own code, unmodified:
  * np_sims/ufunc.c, compiled by oracle/Makefile into oracle/_ref/ (gufuncs hamming,
    hamming_top_10, hamming_top_cand);
  * np_sims/lsh.py and np_sims/hamming.py, exec'd from /root/reference.  `import np_sims`
    itself fails on NumPy >= 2 (np_sims/hamming.py:12, `np.uint64(-1)` raises OverflowError),
    so the package object is assembled here from the same files, with that one literal
    replaced IN MEMORY by its intended value 0xFFFFFFFFFFFFFFFF;
  * test/test_hamming_top_n.py and test/test_hamming.py, imported for their known-answer
    tables.

The reference cannot travel to the GPU box, so only its INPUT/OUTPUT VECTORS are committed
(compressed .npz; the 33 MB KAT #23 is mostly zeros and shrinks to a few KB).  No reference
source is copied.  Seeds for the random cases are recorded next to a checksum of the
regenerated inputs, so a drifting RNG is detected rather than silently accepted.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from tests.golden.cases import RANDOM_TIE_CASES, random_tie_case, sha  # noqa: E402


def load_reference():
    uf = oracle.ref_ufuncs("")
    assert uf is not None, "run `make -C oracle ref` first"
    pkg = types.ModuleType("np_sims")
    pkg.__path__ = []  # mark as package
    sys.modules["np_sims"] = pkg
    sys.modules["np_sims.ufuncs"] = uf
    pkg.ufuncs = uf
    pkg.num_unshared_bits = uf.num_unshared_bits
    pkg.hamming_c = uf.hamming
    pkg.hamming_top_10 = uf.hamming_top_10
    pkg.hamming_top_cand = uf.hamming_top_cand

    ham = types.ModuleType("np_sims.hamming")
    src = open(os.path.join(REF, "np_sims/hamming.py")).read()
    assert "np.uint64(-1)" in src
    exec(compile(src.replace("np.uint64(-1)", "np.uint64(0xFFFFFFFFFFFFFFFF)"), "np_sims/hamming.py", "exec"), ham.__dict__)
    sys.modules["np_sims.hamming"] = ham
    pkg.hamming = ham
    pkg.hamming_naive = ham.hamming_naive

    spec = importlib.util.spec_from_file_location("np_sims.lsh", os.path.join(REF, "np_sims/lsh.py"))
    lsh = importlib.util.module_from_spec(spec)
    sys.modules["np_sims.lsh"] = lsh
    spec.loader.exec_module(lsh)
    pkg.lsh = lsh
    return pkg, uf, lsh, ham


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    pkg, uf, lsh, ham = load_reference()

    # ---- 1. the reference's top-10 KAT table (test/test_hamming_top_n.py:30-166)
    kat = load_by_path("ref_test_hamming_top_n", os.path.join(REF, "test/test_hamming_top_n.py"))
    out = {}
    for i, (hashes, query, expected) in enumerate(kat.hamming_tests):
        h = np.array(hashes, dtype=np.uint64)
        q = np.array(query, dtype=np.uint64)
        got = uf.hamming_top_10(h, q)
        assert set(got.tolist()) == set(int(x) for x in expected), i   # the reference passes its own test
        out[f"hashes_{i}"] = h
        out[f"query_{i}"] = q
        out[f"expected_{i}"] = np.array([int(x) for x in expected], dtype=np.uint64)
        out[f"ref_top10_{i}"] = got                       # slot order, as returned
        out[f"ref_dist_{i}"] = uf.hamming(h, q)
        if h.shape[0] <= 31000:
            out[f"ref_cand_{i}"] = uf.hamming_top_cand(h, q)
    out["num_cases"] = np.array(len(kat.hamming_tests))
    np.savez_compressed(os.path.join(HERE, "kat_top10.npz"), **out)

    # ---- 2. distance / popcount KATs (test/test_hamming.py:7-26)
    kh = load_by_path("ref_test_hamming", os.path.join(REF, "test/test_hamming.py"))
    out = {}
    for i, (bits, expected) in enumerate(kh.bit_tests):
        arr = np.array(bits, dtype=np.uint64)
        out[f"bits_{i}"] = arr.copy()
        got = ham.bit_count64(arr.copy())
        assert (got == expected).all()
        out[f"bits_expected_{i}"] = np.asarray(got, dtype=np.uint64)
    for i, (hashes, query, expected) in enumerate(kh.hamming_tests):
        h, q = np.array(hashes, dtype=np.uint64), np.array(query, dtype=np.uint64)
        assert (uf.hamming(h, q) == expected).all() and (ham.hamming_naive(h, q) == expected).all()
        out[f"hashes_{i}"], out[f"query_{i}"] = h, q
        out[f"expected_{i}"] = np.array(expected, dtype=np.uint64)
    out["num_bit_cases"] = np.array(len(kh.bit_tests))
    out["num_hamming_cases"] = np.array(len(kh.hamming_tests))
    np.savez_compressed(os.path.join(HERE, "kat_hamming.npz"), **out)

    # ---- 3. BASELINE config C1: glove_sample.npy, 640 bits, self-queries, k=10
    X = np.load(os.path.join(REF, "test/glove_sample.npy"))
    assert X.shape == (1019, 300) and X.dtype == np.float32
    np.random.seed(0)
    W = lsh.create_projections(640, 300)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        H = lsh.index(X, W)
    assert H.shape == (1019, 10) and H.dtype == np.uint64
    top10 = np.stack([uf.hamming_top_10(H, H[i].copy()) for i in range(len(H))])
    q_ids = [0, 1, 500, 774, 1018]                       # 774 = "king" (test/test_partition.py:165)
    cand = np.stack([uf.hamming_top_cand(H, H[i].copy()) for i in q_ids])
    dist = np.stack([uf.hamming(H, H[i].copy()) for i in q_ids])
    # lsh.query end to end on the float vectors (np_sims/lsh.py:70-74)
    e2e = np.stack([lsh.query(X[i], H, W)[0] for i in q_ids])
    # argsort variant (np_sims/lsh.py:47-56) — ids are sort-implementation dependent on ties, store dists only
    arg_d = np.stack([lsh.query_with_hamming_then_slow_argsort(X[i], H, W, n=10)[1] for i in q_ids])
    # two-phase (np_sims/lsh.py:84-91), front = 128 bits
    front = lsh.front_hashes(H, 128)
    two = np.stack([lsh.query_two_phase(X[i], front, H, W)[0] for i in q_ids])
    np.savez_compressed(
        os.path.join(HERE, "c1_glove.npz"),
        X=X, H=H, W_sha=np.array(sha(W)), W_first=W[0, :8].copy(), ref_top10=top10,
        q_ids=np.array(q_ids), ref_cand=cand, ref_dist=dist, ref_query=e2e,
        ref_argsort_dists=arg_d, ref_two_phase=two,
    )

    # ---- 4. seeded random signatures with heavy ties: reference outputs in slot order
    out = {}
    cases = []
    for ci, (n, w, bits_kept, seed) in enumerate(RANDOM_TIE_CASES):
        h, q = random_tie_case(n, w, bits_kept, seed)
        cases.append((n, w, bits_kept, seed))
        out[f"in_sha_{ci}"] = np.array(sha(h) + sha(q))
        out[f"ref_top10_{ci}"] = uf.hamming_top_10(h, q)
        out[f"ref_cand_{ci}"] = uf.hamming_top_cand(h, q)
        out[f"ref_dist_{ci}"] = uf.hamming(h, q).astype(np.uint16)
    out["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "random_ties.npz"), **out)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
