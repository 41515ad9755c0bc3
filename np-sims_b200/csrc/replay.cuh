#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nps {
// dist: [nq, N] uint16 (dist_bytes=2) or uint32 (4).  out_rows [nq, k] in the reference's slot
// order, UINT64_MAX where a slot was never filled; out_scores (nullable) the slot distances.
cudaError_t launch_replay_queue(const void* dist, int dist_bytes, unsigned long long N, uint32_t nq, uint32_t k,
                                uint64_t* out_rows, uint64_t* out_scores, cudaStream_t stream,
                                const uint32_t* gate = nullptr);   // gate: no-op unless *gate != 0
}  // namespace nps
