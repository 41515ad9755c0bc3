#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nps {
// K1a: float64 CUDA-core projection, sign, pack.  out [n_rows, n_proj/64] (nullable), dots
// [n_rows, n_proj] float64 (nullable, for tolerance analysis).
cudaError_t launch_project_f64(const void* X, int x_is_f32, unsigned long long n_rows, uint32_t dims,
                               const double* W, uint32_t n_proj, uint64_t* out, double* dots, cudaStream_t stream);
}  // namespace nps
