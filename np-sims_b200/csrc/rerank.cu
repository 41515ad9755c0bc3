// rerank.cu — K6: second phase of the two-phase search (np_sims/lsh.py:84-91): for every query,
// the exact full-width Hamming distance to each of its candidate rows (the reference gathers
// `hashes[candidates]` and runs hamming_top_10 on the copy, lsh.py:88-90) and the n best of them,
// without materialising the gathered table: one CTA per query reads the candidate rows in place.
//
// Order: ascending (distance, row id) — the canonical rule of this library.  Candidate ids of
// UINT64_MAX (padding of a short first phase) or beyond the table are ignored; a candidate listed more than
// once counts once.
#include "common.cuh"
#include "rerank.cuh"

namespace nps {

constexpr int kRerankThreads = 256;

__global__ void __launch_bounds__(kRerankThreads) rerank_kernel(const uint64_t* __restrict__ H, unsigned long long N,
                                                                uint32_t W, const uint64_t* __restrict__ Q,
                                                                const uint64_t* __restrict__ cand, uint32_t C,
                                                                uint32_t P, uint32_t k, uint64_t* __restrict__ out_rows,
                                                                uint64_t* __restrict__ out_dists) {
    extern __shared__ __align__(16) uint64_t sm[];
    uint64_t* keys = sm;          // [P] (P = power of two >= C)
    uint64_t* qs = sm + P;        // [W]
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    for (uint32_t w = tid; w < W; w += kRerankThreads) qs[w] = Q[(size_t)q * W + w];
    __syncthreads();
    for (uint32_t i = tid; i < P; i += kRerankThreads) {
        uint64_t key = kNone;
        if (i < C) {
            const uint64_t row = cand[(size_t)q * C + i];
            if (row < N) {
                const uint64_t* h = H + row * W;
                uint32_t d = 0;
                if ((W & 1u) == 0u) {   // rows are 16-byte aligned: one 128-bit load per two words
                    for (uint32_t w = 0; w < W; w += 2) {
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(h + w);
                        d += __popcll(v.x ^ qs[w]) + __popcll(v.y ^ qs[w + 1]);
                    }
                } else {
                    for (uint32_t w = 0; w < W; ++w) d += __popcll(__ldg(h + w) ^ qs[w]);
                }
                key = ((uint64_t)d << kGlobalRowBits) | row;
            }
        }
        keys[i] = key;
    }
    __syncthreads();
    // bitonic sort, ascending (pads sort last)
    auto sort_keys = [&]() {
        for (uint32_t size = 2; size <= P; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t i = tid; i < (P >> 1); i += kRerankThreads) {
                    const uint32_t pos = 2 * i - (i & (stride - 1));
                    const uint64_t a = keys[pos], b = keys[pos + stride];
                    const bool up = (pos & size) == 0;
                    if ((a > b) == up) {
                        keys[pos] = b;
                        keys[pos + stride] = a;
                    }
                }
                __syncthreads();
            }
        }
    };
    sort_keys();
    // A candidate listed more than once (arbitrary caller lists; hamming_top_n never emits duplicates):
    // equal keys are now neighbours — blank all but the first of each run and sort again, so the result
    // is compact and the (k+1)-th distinct candidate moves up.
    constexpr uint32_t kPerThread = 4096 / kRerankThreads;
    bool dup[kPerThread];
    bool any = false;
#pragma unroll
    for (uint32_t j = 0; j < kPerThread; ++j) {
        const uint32_t i = tid + j * kRerankThreads;
        dup[j] = i > 0 && i < P && keys[i] != kNone && keys[i] == keys[i - 1];
        any |= dup[j];
    }
    if (__syncthreads_or(any)) {
#pragma unroll
        for (uint32_t j = 0; j < kPerThread; ++j)
            if (dup[j]) keys[tid + j * kRerankThreads] = kNone;
        __syncthreads();
        sort_keys();
    }
    for (uint32_t i = tid; i < k; i += kRerankThreads) {
        const uint64_t key = i < P ? keys[i] : kNone;
        const size_t o = (size_t)q * k + i;
        out_rows[o] = key == kNone ? kNone : (key & ((1ull << kGlobalRowBits) - 1ull));
        if (out_dists) out_dists[o] = key == kNone ? kNone : (key >> kGlobalRowBits);
    }
}

cudaError_t launch_rerank(const uint64_t* H, unsigned long long N, uint32_t W, const uint64_t* Q, uint32_t nq,
                          const uint64_t* cand, uint32_t C, uint32_t k, uint64_t* out_rows, uint64_t* out_dists,
                          cudaStream_t stream) {
    if (nq == 0 || k == 0) return cudaSuccess;
    uint32_t P = 2;
    while (P < C) P <<= 1;
    const size_t smem = ((size_t)P + W) * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rerank_kernel<<<nq, kRerankThreads, smem, stream>>>(H, N, W, Q, cand, C, P, k, out_rows, out_dists);
    count_launch();
    return cudaGetLastError();
}

}  // namespace nps
