// merge.cu — K3 block-merge / K4 shard-merge: reduce L sorted candidate lists per query to
// the k best.  Replaces the final state of the reference's queue (np_sims/ufunc.c:150-195)
// and `np.argsort(...)[:n]` (np_sims/lsh.py:54).
//
// Keys are 64-bit  dist << 40 | global_row  (unique per row), padding = UINT64_MAX, each
// input list ascending.  A CTA takes a group of G lists of one query into shared memory and
// places every element directly at its output rank
//       rank(e) = pos_in_own_list + sum_{other lists} lower_bound(list, e)
// (keys are unique, so ranks are a permutation; pads rank >= #real keys).  No sort network,
// no atomics; ceil(log_G L) passes.  The last pass decodes keys to (row id, distance).
#include "common.cuh"
#include "merge.cuh"

namespace nps {

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* a, uint32_t n, uint64_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// in : [Q, lists_in, k]   out_keys: [Q, lists_out, k]  (lists_out = ceil(lists_in / G))
// final pass (out_rows != nullptr, lists_out == 1): rows [Q,k] u64, dists [Q,k] u64 (may be null)
__global__ void __launch_bounds__(256) merge_lists_kernel(const uint64_t* __restrict__ in, uint32_t Q, uint32_t lists_in,
                                                          size_t q_stride, size_t list_stride,
                                                          uint32_t k, uint32_t G, uint64_t* __restrict__ out_keys,
                                                          uint64_t* __restrict__ out_rows,
                                                          uint64_t* __restrict__ out_dists,
                                                          const uint32_t* __restrict__ gate) {
    if (gate != nullptr && *gate == 0u) return;   // gated repair launch (api.cu), nothing to repair
    extern __shared__ __align__(16) uint64_t skeys[];
    const uint32_t grp = blockIdx.y, lists_out = gridDim.y;
    const uint32_t g0 = grp * G;
    const uint32_t g = (lists_in - g0) < G ? (lists_in - g0) : G;
    const uint32_t n = g * k;
    // queries on x, grid-stride (gated launches use a small grid so that a no-op costs ~2 us)
    for (uint32_t q = blockIdx.x; q < Q; q += gridDim.x) {
    const uint64_t* src = in + (size_t)q * q_stride + (size_t)g0 * list_stride;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t a = i / k;
        skeys[i] = src[(size_t)a * list_stride + (i - a * k)];
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
        const uint32_t a = e / k, i = e - a * k;
        const uint64_t key = skeys[e];
        uint32_t rank = i;
        for (uint32_t b = 0; b < g && rank < k; ++b)
            if (b != a) rank += lower_bound_u64(skeys + b * k, k, key);
        if (rank < k) {
            if (out_rows) {
                const size_t o = (size_t)q * k + rank;
                out_rows[o] = key == kNone ? kNone : (key & ((1ull << kGlobalRowBits) - 1ull));
                if (out_dists) out_dists[o] = key == kNone ? kNone : (key >> kGlobalRowBits);
            } else {
                out_keys[((size_t)q * lists_out + grp) * k + rank] = key;
            }
        }
    }
    }
}

// Few short lists (the single-query / small-batch scan: ~148 lists of k <= 64 keys): the rank merge
// above costs O(n * lists * log k) shared-memory probes on ONE CTA (0.13 ms for 148 x 10 keys).
// Here the CTA keeps the n <= 4096 keys in shared memory and extracts the k smallest by repeated
// "smallest key greater than the previous winner" (keys are unique): k rounds of a strided scan +
// shuffle/shared-memory min — ~150 cycles per round.
constexpr uint32_t kSmallMergeKeys = 4096, kSmallMergeK = 64;
__global__ void __launch_bounds__(256) merge_small_kernel(const uint64_t* __restrict__ in, uint32_t Q, uint32_t lists,
                                                          size_t q_stride, size_t list_stride, uint32_t k,
                                                          uint64_t* __restrict__ out_keys,
                                                          uint64_t* __restrict__ out_rows,
                                                          uint64_t* __restrict__ out_dists,
                                                          const uint32_t* __restrict__ gate) {
    if (gate != nullptr && *gate == 0u) return;
    __shared__ uint64_t skeys[kSmallMergeKeys];
    __shared__ uint64_t wmin[8];
    const uint32_t n = lists * k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t q = blockIdx.x; q < Q; q += gridDim.x) {
    const uint64_t* src = in + (size_t)q * q_stride;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += 256) {
        const uint32_t a = i / k;
        skeys[i] = src[(size_t)a * list_stride + (i - a * k)];
    }
    __syncthreads();
    uint64_t prev = 0;
    bool first = true;
    for (uint32_t r = 0; r < k; ++r) {
        uint64_t best = kNone;
        for (uint32_t i = tid; i < n; i += 256) {
            const uint64_t v = skeys[i];
            if ((first || v > prev) && v < best) best = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if (lane == 0) wmin[warp] = best;
        __syncthreads();
        best = wmin[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) best = wmin[w] < best ? wmin[w] : best;
        __syncthreads();
        if (tid == 0) {
            const size_t o = (size_t)q * k + r;
            if (out_rows) {
                out_rows[o] = best == kNone ? kNone : (best & ((1ull << kGlobalRowBits) - 1ull));
                if (out_dists) out_dists[o] = best == kNone ? kNone : (best >> kGlobalRowBits);
            } else {
                out_keys[o] = best;
            }
        }
        prev = best;          // kNone from here on: nothing is > kNone, every later slot is a pad
        first = false;
    }
    }
}

// Lists merged per CTA and pass.  As many as fit in shared memory when there are plenty of queries to fill
// the GPU; with few queries (the single-query scan: 148 lists, 1 CTA per group) a CTA that ranks 4096 keys
// against 127 other lists is the whole kernel (0.39 ms at k = 32), so the groups shrink to ~sqrt(lists):
// more CTAs in the first pass, a short second one.
uint32_t merge_group_size(uint32_t k, uint32_t Q, uint32_t lists) {
    uint32_t G = kMergeSmemKeys / (k ? k : 1);
    if (G < 2) G = 2;
    if ((uint64_t)Q * ((lists + G - 1) / G) < 148u) {
        uint32_t r = 2;
        while (r * r < lists) ++r;
        if (r < G) G = r;
    }
    return G;
}

size_t merge_scratch_keys(uint32_t Q, uint32_t lists, uint32_t k) {
    // ping-pong buffer for intermediate passes: the first pass shrinks lists by G
    if (k <= kSmallMergeK && (uint64_t)lists * k <= kSmallMergeKeys) return 0;   // merge_small_kernel: one pass
    const uint32_t G = merge_group_size(k, Q, lists);
    if (lists <= G) return 0;
    const uint32_t l1 = (lists + G - 1) / G;
    const uint32_t l2 = (l1 + G - 1) / G;
    return (size_t)Q * k * ((size_t)l1 + (l1 > G ? l2 : 0));
}

cudaError_t launch_merge(const uint64_t* keys, uint32_t Q, uint32_t lists, uint32_t k, int list_major,
                         uint64_t* scratch, uint64_t* out_rows, uint64_t* out_dists, uint64_t* out_keys,
                         cudaStream_t stream, const uint32_t* gate) {
    if (Q == 0 || k == 0) return cudaSuccess;
    if (k <= kSmallMergeK && (uint64_t)lists * k <= kSmallMergeKeys && lists > 1) {
        merge_small_kernel<<<gate ? (Q < 592u ? Q : 592u) : Q, 256, 0, stream>>>(keys, Q, lists, list_major ? (size_t)k : (size_t)lists * k,
                                                  list_major ? (size_t)Q * k : (size_t)k, k, out_keys,
                                                  out_keys ? nullptr : out_rows, out_keys ? nullptr : out_dists, gate);
        count_launch();
        return cudaGetLastError();
    }
    const uint32_t G = merge_group_size(k, Q, lists);
    const uint32_t gx = gate ? (Q < 592u ? Q : 592u) : Q;   // gated: small grid-stride grid
    const size_t smem = (size_t)(G < lists ? G : lists) * k * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kMergeSmemKeys * 2 * sizeof(uint64_t)));
    if (e != cudaSuccess) return e;
    const uint64_t* cur = keys;
    uint32_t cur_lists = lists;
    size_t q_stride = list_major ? (size_t)k : (size_t)lists * k;
    size_t list_stride = list_major ? (size_t)Q * k : (size_t)k;
    uint64_t* bufA = scratch;
    uint64_t* bufB = scratch ? scratch + (size_t)Q * k * ((lists + G - 1) / G) : nullptr;
    int flip = 0;
    while (cur_lists > G) {
        const uint32_t nxt = (cur_lists + G - 1) / G;
        uint64_t* dst = flip ? bufB : bufA;
        merge_lists_kernel<<<dim3(gx, nxt), 256, (size_t)G * k * 8, stream>>>(cur, Q, cur_lists, q_stride, list_stride, k,
                                                                             G, dst, nullptr, nullptr, gate);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        cur = dst;
        cur_lists = nxt;
        q_stride = (size_t)nxt * k;
        list_stride = k;
        flip ^= 1;
    }
    // last pass: decoded ids/distances, or (out_keys) the merged key list itself
    merge_lists_kernel<<<dim3(gx, 1), 256, smem, stream>>>(cur, Q, cur_lists, q_stride, list_stride, k, G, out_keys,
                                                          out_keys ? nullptr : out_rows, out_keys ? nullptr : out_dists,
                                                          gate);
    count_launch();
    return cudaGetLastError();
}

}  // namespace nps
