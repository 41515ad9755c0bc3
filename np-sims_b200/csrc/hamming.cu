// hamming.cu — K5: full Hamming distances (replaces gufunc `hamming`, np_sims/ufunc.c:74-106,
// and np_sims/hamming.py:41-61) and the elementwise popcount(a^b) (ufunc.c:41-66).
//
// out[q, i] = sum_w popc(H[i,w] ^ Q[q,w]).  HBM-bound: reads N*W*8 bytes once per batch of
// queries, writes Q*N*sizeof(out).  A CTA copies a tile of rows into shared memory with
// fully coalesced 16-byte loads (row stride padded to an odd number of words, so the
// row-per-thread reads that follow are bank-conflict free), then every thread scores its row
// against each staged query; the [q, row] stores of a warp are contiguous.
#include "common.cuh"
#include "hamming.cuh"

namespace nps {

constexpr int kDistThreads = 256;
constexpr int kDistQBatch = 16;

template <typename OutT>
__global__ void __launch_bounds__(kDistThreads) hamming_dist_kernel(const uint64_t* __restrict__ H,
                                                                    unsigned long long N, uint32_t W,
                                                                    const uint64_t* __restrict__ Q, uint32_t nq,
                                                                    OutT* __restrict__ out, uint32_t tile_rows,
                                                                    const uint32_t* __restrict__ gate) {
    extern __shared__ __align__(16) uint64_t sm[];
    if (gate != nullptr && *gate == 0u) return;   // gated fallback launch, nothing to redo
    const uint32_t stride = W | 1u;
    uint64_t* rows_s = sm;                                // [tile_rows][stride]
    uint64_t* q_s = sm + (size_t)tile_rows * stride;      // [kDistQBatch][W]
    const int tid = threadIdx.x;
    const unsigned long long num_tiles = (N + tile_rows - 1) / tile_rows;
    for (unsigned long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const unsigned long long row0 = tile * tile_rows;
        const uint32_t rows = (N - row0) < tile_rows ? (uint32_t)(N - row0) : tile_rows;
        const uint64_t* src = H + row0 * W;
        const uint32_t words = rows * W;
        __syncthreads();
        for (uint32_t i = tid; i < words; i += kDistThreads) {
            const uint32_t r = i / W, w = i - r * W;
            rows_s[r * stride + w] = __ldg(src + i);
        }
        const bool active = (uint32_t)tid < rows;
        const uint64_t* my = rows_s + (size_t)(active ? tid : 0) * stride;
        for (uint32_t qb = 0; qb < nq; qb += kDistQBatch) {
            const uint32_t nb = (nq - qb) < kDistQBatch ? (nq - qb) : kDistQBatch;
            __syncthreads();
            for (uint32_t i = tid; i < nb * W; i += kDistThreads) q_s[i] = __ldg(Q + (size_t)qb * W + i);
            __syncthreads();
            if (active) {
                for (uint32_t q = 0; q < nb; ++q) {
                    const uint64_t* qq = q_s + q * W;
                    uint32_t d = 0;
                    for (uint32_t w = 0; w < W; ++w) d += __popcll(my[w] ^ qq[w]);
                    out[(size_t)(qb + q) * N + row0 + tid] = (OutT)d;
                }
            }
        }
    }
}

uint32_t hamming_tile_rows(uint32_t W) {
    const uint32_t stride = W | 1u;
    uint32_t rows = (96u * 1024u) / (stride * 8u);
    if (rows > (uint32_t)kDistThreads) rows = kDistThreads;
    return rows;   // 0 => width unsupported
}

template <typename OutT>
static cudaError_t launch_dist_t(const uint64_t* H, unsigned long long N, uint32_t W, const uint64_t* Q, uint32_t nq,
                                 void* out, int sm_count, cudaStream_t stream, const uint32_t* gate) {
    const uint32_t tile_rows = hamming_tile_rows(W);
    const size_t smem = ((size_t)tile_rows * (W | 1u) + (size_t)kDistQBatch * W) * 8;
    auto kern = hamming_dist_kernel<OutT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    unsigned long long tiles = (N + tile_rows - 1) / tile_rows;
    unsigned long long grid = tiles < (unsigned long long)sm_count * 8 ? tiles : (unsigned long long)sm_count * 8;
    kern<<<(unsigned)grid, kDistThreads, smem, stream>>>(H, N, W, Q, nq, (OutT*)out, tile_rows, gate);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_hamming_dist(const uint64_t* H, unsigned long long N, uint32_t W, const uint64_t* Q, uint32_t nq,
                                void* out, int out_bytes, int sm_count, cudaStream_t stream, const uint32_t* gate) {
    if (N == 0 || nq == 0) return cudaSuccess;
    switch (out_bytes) {
        case 2: return launch_dist_t<uint16_t>(H, N, W, Q, nq, out, sm_count, stream, gate);
        case 4: return launch_dist_t<uint32_t>(H, N, W, Q, nq, out, sm_count, stream, gate);
        case 8: return launch_dist_t<uint64_t>(H, N, W, Q, nq, out, sm_count, stream, gate);
        default: return cudaErrorInvalidValue;
    }
}

__global__ void unshared_bits_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                     unsigned long long n, unsigned long long a_mod, unsigned long long b_mod,
                                     uint8_t* __restrict__ out) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = (uint8_t)__popcll(a[i % a_mod] ^ b[i % b_mod]);
}

cudaError_t launch_unshared_bits(const uint64_t* a, unsigned long long na, const uint64_t* b, unsigned long long nb,
                                 unsigned long long n, uint8_t* out, int sm_count, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n + 255) / 256;
    if (blocks > (unsigned long long)sm_count * 16) blocks = (unsigned long long)sm_count * 16;
    unshared_bits_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a, b, n, na, nb, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace nps
