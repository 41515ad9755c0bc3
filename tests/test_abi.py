"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/npsims_b200.h declares, validates arguments without a GPU, and the Python surface
mirrors the reference's names."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "npsims_b200.h")
LIB = os.path.join(ROOT, "np-sims_b200", "lib", "libnpsims_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run `python np-sims_b200/build.py`"
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_header():
    from np_sims import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    assert _lib.raw("nps_version")() >= 100


def test_workspace_queries_need_no_gpu():
    from np_sims import _lib
    f = _lib.raw("nps_hamming_top_n_workspace_bytes")
    assert f(400_000, 10, 10_000, 10) > 0
    assert f(100_000_000 // 8, 16, 1, 100) > 0
    assert f(1000, 0, 1, 10) == 0            # words == 0
    assert f(1000, 10, 1, 5000) == 0         # k > NPS_MAX_K
    assert "k=5000" in _lib.last_error()
    assert _lib.raw("nps_top_n_merge_workspace_bytes")(100, 8, 10) == 0
    g = _lib.raw("nps_hamming_top_n_reference_workspace_bytes")
    assert g(1000, 10, 2, 10) >= 2 * 1000 * 2              # dense fallback rows are always included
    assert g(400_000, 10, 1, 10) < 4 << 20                 # K2R blocks are small next to the table
    assert g(25_000_000, 16, 1, 1000) >= 25_000_000 * 2
    assert g(1000, 10, 2, 0) == 0 and g(1000, 0, 2, 10) == 0


def test_reference_surface_names():
    import np_sims
    import np_sims.lsh as lsh
    import np_sims.hamming as ham
    for name in ("num_unshared_bits", "hamming_c", "hamming_top_10", "hamming_top_cand", "hamming_naive"):
        assert callable(getattr(np_sims, name))            # reference np_sims/__init__.py:1-6
    for name in ("random_projection", "create_projections", "index_one", "index",
                 "query_with_hamming_then_slow_argsort", "query_pure_python", "query_with_hamming_top_n",
                 "front_hashes", "query_with_hamming_top_n_two_phase", "query", "query_two_phase"):
        assert callable(getattr(lsh, name))                # reference np_sims/lsh.py:5-95
    assert lsh.query is lsh.query_with_hamming_top_n and lsh.query_two_phase is lsh.query_with_hamming_top_n_two_phase
    assert callable(ham.bit_count64) and callable(ham.hamming_naive)


@pytest.mark.parametrize("seed", [0, 1000])
@pytest.mark.parametrize("num_projections", [64, 128, 256])
@pytest.mark.parametrize("num_dims", [50, 100, 300])
def test_projections_normalized(seed, num_projections, num_dims):
    # mirrors reference test/test_lsh.py:34-42; and the stream equals the oracle's restatement
    import np_sims.lsh as lsh
    from oracle import oracle
    np.random.seed(seed)
    projections = lsh.create_projections(num_projections, dims=num_dims)
    for proj in projections:
        assert pytest.approx(np.linalg.norm(proj)) == 1.0
    np.random.seed(seed)
    assert np.array_equal(np.asarray(projections), oracle.create_projections(num_projections, num_dims))


def test_errors_without_gpu_match_reference_types():
    import torch
    import np_sims
    import np_sims.lsh as lsh
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    h = np.zeros((4, 2), dtype=np.uint64)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        np_sims.hamming_top_10(h, h[0])
    with pytest.raises(ValueError, match="multiple of 64"):
        lsh.index(np.zeros((3, 8)), np.zeros((100, 8)))


def test_hash_split_buffer_size_and_path_switch_need_no_gpu():
    from np_sims import _lib
    size = _lib.raw("nps_project_split_bytes")(640, 300)
    tf32 = (2 * 640 * 300 * 4 + 255) // 256 * 256              # float32 [2, B, D]
    bf16 = (2 * 640 * 304 * 2 + 255) // 256 * 256              # bfloat16 [2, B, D rounded up to 8]
    assert size == tf32 + bf16 + 4 * 5 * 4                     # + float32 prefix norms, one per 16 dims of 5 K-blocks
    assert _lib.raw("nps_set_hash_path")(7) != 0 and "hash path" in _lib.last_error()
    for path in (_lib.HASH_TF32, _lib.HASH_BF16, _lib.HASH_AUTO):
        assert _lib.raw("nps_set_hash_path")(path) == 0
    assert _lib.raw("nps_set_scan_path")(9) != 0
