#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nps {
// K1a: float64 CUDA-core projection, sign, pack.  out [n_rows, n_proj/64] (nullable), dots
// [n_rows, n_proj] float64 (nullable, for tolerance analysis).
// gate (nullable): the launch is a no-op unless *gate > gate_above (fix-up list overflow repair).
cudaError_t launch_project_f64(const void* X, int x_is_f32, unsigned long long n_rows, uint32_t dims,
                               const double* W, uint32_t n_proj, uint64_t* out, double* dots, cudaStream_t stream,
                               const unsigned int* gate = nullptr, unsigned int gate_above = 0);

// K1T: tensor-core projection (project_tc.cu).  float32 vectors only; dims % 4 == 0.
bool project_tc_applicable(unsigned long long N, uint32_t D, uint32_t B);
size_t project_tc_workspace_bytes(unsigned long long N, uint32_t B, uint32_t* fix_cap_out);
// Wsplit: project_split_bytes(B, D) bytes from launch_split_projections — the TF32 high / low parts of
// W64 as float32 [2, B, D], then the BF16 high / low parts as bfloat16 [2, B, Dp].
size_t project_split_bytes(uint32_t B, uint32_t D);
cudaError_t launch_split_projections(const double* W64, uint32_t B, uint32_t D, float* out, cudaStream_t stream);
cudaError_t launch_project_tc(const float* X, unsigned long long N, uint32_t D, const float* Wsplit, const double* W64,
                              uint32_t B, float w_norm_max, uint64_t* out, void* ws, int sm_count, int smem_optin,
                              cudaStream_t stream, bool use_bf16 = true);
}  // namespace nps
