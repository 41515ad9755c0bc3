// project_f64.cu — K1a: exact-arithmetic hashing  sign(X . W^T) -> packed uint64
// (replaces lsh.index_one / lsh.index, np_sims/lsh.py:31-44).
//
// The reference computes the projections in float64 (np.dot of float64 projections with the
// vector, lsh.py:32) and sets bit j of the signature when dot_j >= 0 (bit j -> word j/64,
// position j%64; NaN -> 0, exact zero -> 1).  This kernel does the same arithmetic on the
// FP64 pipe of the CUDA cores (DFMA), so its bits agree with the reference except where the
// float64 summation ORDER matters (|x.w| ~ 1e-16).  It is the parity mode of the hashing step;
// the throughput mode is the tcgen05 kernel in project_tc.cu.
//
// Tiling: a CTA produces 64 rows x 64 projections = one uint64 word for each of 64 rows.
// X and W tiles of 16 columns are staged in shared memory as float64 (X converted on load);
// 256 threads hold 4x4 accumulators each.
#include "common.cuh"
#include "project.cuh"

namespace nps {

constexpr int kPM = 64, kPN = 64, kPK = 16;

// one 64-row x 64-projection tile (tile_x, tile_y)
template <typename XT>
__device__ __forceinline__ void project_f64_tile(const XT* __restrict__ X, unsigned long long n_rows, uint32_t dims,
                                                 const double* __restrict__ W, uint32_t n_proj,
                                                 uint64_t* __restrict__ out, double* __restrict__ dots,
                                                 unsigned long long tile_x, uint32_t tile_y) {
    __shared__ double xs[kPK][kPM + 1];
    __shared__ double ws[kPK][kPN + 1];
    __shared__ unsigned long long word_s[kPM];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 4 x 4 outputs each
    const unsigned long long row0 = tile_x * kPM;
    const uint32_t proj0 = tile_y * kPN;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    if (tid < kPM) word_s[tid] = 0ull;

    for (uint32_t k0 = 0; k0 < dims; k0 += kPK) {
        __syncthreads();
        for (int i = tid; i < kPM * kPK; i += 256) {
            const int r = i / kPK, c = i % kPK;
            const unsigned long long row = row0 + r;
            xs[c][r] = (row < n_rows && k0 + c < dims) ? (double)X[row * dims + k0 + c] : 0.0;
        }
        for (int i = tid; i < kPN * kPK; i += 256) {
            const int r = i / kPK, c = i % kPK;
            const uint32_t pj = proj0 + r;
            ws[c][r] = (pj < n_proj && k0 + c < dims) ? W[(size_t)pj * dims + k0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < kPK; ++c) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[c][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = ws[c][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned long long bits = 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (acc[i][j] >= 0.0) bits |= 1ull << (tx * 4 + j);
            if (dots) {
                const unsigned long long row = row0 + ty * 4 + i;
                const uint32_t pj = proj0 + tx * 4 + j;
                if (row < n_rows && pj < n_proj) dots[row * n_proj + pj] = acc[i][j];
            }
        }
        atomicOr(&word_s[ty * 4 + i], bits);
    }
    __syncthreads();
    if (tid < kPM && row0 + tid < n_rows && out)
        out[(row0 + tid) * (n_proj / 64) + tile_y] = word_s[tid];
}

template <typename XT>
__global__ void __launch_bounds__(256) project_f64_kernel(const XT* __restrict__ X, unsigned long long n_rows,
                                                          uint32_t dims, const double* __restrict__ W,
                                                          uint32_t n_proj, uint64_t* __restrict__ out,
                                                          double* __restrict__ dots) {
    project_f64_tile<XT>(X, n_rows, dims, W, n_proj, out, dots, blockIdx.x, blockIdx.y);
}

// Gated repair launch (project_tc.cu): a small persistent grid that does nothing unless the
// tensor-core pass overflowed its fix-up list (*gate > gate_above), then redoes every tile.
__global__ void __launch_bounds__(256) project_f64_gated_kernel(const float* __restrict__ X, unsigned long long n_rows,
                                                                uint32_t dims, const double* __restrict__ W,
                                                                uint32_t n_proj, uint64_t* __restrict__ out,
                                                                const unsigned int* __restrict__ gate,
                                                                unsigned int gate_above) {
    if (*gate <= gate_above) return;
    const unsigned long long tiles_x = (n_rows + kPM - 1) / kPM;
    const uint32_t tiles_y = (n_proj + kPN - 1) / kPN;
    for (unsigned long long t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
        __syncthreads();
        project_f64_tile<float>(X, n_rows, dims, W, n_proj, out, nullptr, t / tiles_y, (uint32_t)(t % tiles_y));
    }
}

cudaError_t launch_project_f64(const void* X, int x_is_f32, unsigned long long n_rows, uint32_t dims,
                               const double* W, uint32_t n_proj, uint64_t* out, double* dots, cudaStream_t stream,
                               const unsigned int* gate, unsigned int gate_above) {
    if (n_rows == 0) return cudaSuccess;
    const unsigned long long gx = (n_rows + kPM - 1) / kPM;
    if (gx > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    if (gate) {
        if (!x_is_f32 || dots) return cudaErrorInvalidValue;
        const unsigned long long tiles = gx * ((n_proj + kPN - 1) / kPN);
        project_f64_gated_kernel<<<(unsigned)(tiles < 2368 ? tiles : 2368), 256, 0, stream>>>(
            (const float*)X, n_rows, dims, W, n_proj, out, gate, gate_above);
        count_launch();
        return cudaGetLastError();
    }
    dim3 grid((unsigned)gx, (n_proj + kPN - 1) / kPN);
    if (x_is_f32)
        project_f64_kernel<float><<<grid, 256, 0, stream>>>((const float*)X, n_rows, dims, W, n_proj, out, dots);
    else
        project_f64_kernel<double><<<grid, 256, 0, stream>>>((const double*)X, n_rows, dims, W, n_proj, out, dots);
    count_launch();
    return cudaGetLastError();
}

}  // namespace nps
