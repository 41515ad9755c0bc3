#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace nps {
constexpr uint32_t kMergeSmemKeys = 4096;   // keys a merge CTA holds (32 KB)
uint32_t merge_group_size(uint32_t k, uint32_t Q, uint32_t lists);
size_t merge_scratch_keys(uint32_t Q, uint32_t lists, uint32_t k);
// keys [Q, lists, k] (or [lists, Q, k] if list_major), each list ascending and UINT64_MAX padded
//   -> out_keys [Q, k] if out_keys != nullptr, else decoded rows/dists [Q, k]
cudaError_t launch_merge(const uint64_t* keys, uint32_t Q, uint32_t lists, uint32_t k, int list_major,
                         uint64_t* scratch, uint64_t* out_rows, uint64_t* out_dists, uint64_t* out_keys,
                         cudaStream_t stream, const uint32_t* gate = nullptr);
}  // namespace nps
