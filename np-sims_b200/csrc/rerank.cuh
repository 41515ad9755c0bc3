#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nps {
// K6: out_rows/out_dists [nq, k] = the k candidates of each query (cand [nq, C] row ids) with the
// smallest full-width distance, ascending (distance, row id); UINT64_MAX padding.
cudaError_t launch_rerank(const uint64_t* H, unsigned long long N, uint32_t W, const uint64_t* Q, uint32_t nq,
                          const uint64_t* cand, uint32_t C, uint32_t k, uint64_t* out_rows, uint64_t* out_dists,
                          cudaStream_t stream);
}  // namespace nps
