#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nps {
uint32_t hamming_tile_rows(uint32_t W);
// out [nq, N] of uint16/uint32/uint64 (out_bytes = 2/4/8)
cudaError_t launch_hamming_dist(const uint64_t* H, unsigned long long N, uint32_t W, const uint64_t* Q, uint32_t nq,
                                void* out, int out_bytes, int sm_count, cudaStream_t stream,
                                const uint32_t* gate = nullptr);   // gate: no-op unless *gate != 0
// out[i] = popc(a[i % na] ^ b[i % nb]), i < n   (scalar/array broadcasting of the elementwise ufunc)
cudaError_t launch_unshared_bits(const uint64_t* a, unsigned long long na, const uint64_t* b, unsigned long long nb,
                                 unsigned long long n, uint8_t* out, int sm_count, cudaStream_t stream);
}  // namespace nps
