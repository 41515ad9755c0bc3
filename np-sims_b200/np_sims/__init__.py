"""np_sims — drop-in B200 implementation of softwaredoug/np-sims' LSH search path.

Same public names as the reference package (/root/reference/np_sims/__init__.py:1-6) plus the
general `hamming_top_n`.  Backed by libnpsims_b200.so (hand-written sm_100a CUDA kernels behind
a C ABI, include/npsims_b200.h); PyTorch is used only to hold device buffers and streams.
There is no CPU fallback."""
from np_sims.ufuncs import num_unshared_bits
from np_sims.ufuncs import hamming as hamming_c
from np_sims.ufuncs import hamming_top_10
from np_sims.ufuncs import hamming_top_cand
from np_sims.ufuncs import hamming_top_n
from np_sims.ufuncs import hamming_rerank

from np_sims.hamming import hamming_naive

__all__ = ["num_unshared_bits", "hamming_c", "hamming_top_10", "hamming_top_cand", "hamming_top_n",
           "hamming_rerank", "hamming_naive"]
