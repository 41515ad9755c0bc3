"""`np_sims.hamming` of the reference (/root/reference/np_sims/hamming.py) on the B200.

The reference's module is the pure-NumPy slow path (SWAR popcount + XOR + sum).  Here the
same two public names run the CUDA kernels; there is no NumPy arithmetic."""
from __future__ import annotations

import numpy as np

from . import ufuncs


def bit_count64(arr):
    """Bits set per uint64 element (reference hamming.py:21-38).  Returns uint64 like the
    reference; unlike the reference it does not clobber its argument (hamming.py:29)."""
    a = np.asarray(arr) if not ufuncs.is_tensor(arr) else arr
    counts = ufuncs.num_unshared_bits(a, np.zeros((), dtype=np.uint64))
    if ufuncs.is_tensor(counts):
        import torch
        return counts.to(torch.int64).view(torch.uint64)
    return counts.astype(np.uint64)


def hamming_naive(hashes, query, count_function=bit_count64):
    """Reference hamming.py:41-61: np.sum(count_function(hashes ^ query), axis=1).
    `count_function` is accepted for signature compatibility; the popcount is fused into the
    distance kernel (K5)."""
    if count_function is not bit_count64:
        raise TypeError("np_sims (B200): only the built-in bit_count64 is supported as count_function")
    return ufuncs.hamming(hashes, query)
