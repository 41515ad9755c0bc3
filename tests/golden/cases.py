"""Seeded input generators shared by make_golden.py (which records the reference's outputs)
and the tests (which regenerate the same inputs and check their checksum)."""
import hashlib

import numpy as np

RANDOM_TIE_CASES = (
    [(3000, w, 64 * w, 100 + w) for w in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 21, 32)]
    + [(5000, 2, 6, 7), (5000, 10, 12, 8), (1500, 16, 5, 9), (9, 4, 256, 10), (10, 3, 192, 11),
       (11, 1, 64, 12), (1, 10, 640, 13), (999, 5, 40, 14), (1000, 5, 40, 15), (1001, 5, 40, 16),
       (20000, 1, 64, 17), (20000, 16, 1024, 18)]
)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def random_tie_case(n: int, w: int, bits_kept: int, seed: int):
    """uint64[n,w] signatures with only `bits_kept` live bit positions (few live bits ->
    many exactly tied distances) and one query; (hashes, query)."""
    rng = np.random.default_rng(seed)
    h = rng.integers(0, 2**64, size=(n, w), dtype=np.uint64)
    if bits_kept < 64 * w:
        mask = np.zeros(w, dtype=np.uint64)
        live = rng.choice(64 * w, size=bits_kept, replace=False)
        for b in live:
            mask[b // 64] |= np.uint64(1) << np.uint64(b % 64)
        h &= mask
    q = h[rng.integers(0, n)].copy() if seed % 2 else rng.integers(0, 2**64, size=w, dtype=np.uint64)
    return h, q
