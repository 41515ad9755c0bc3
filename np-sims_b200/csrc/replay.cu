// replay.cu — reference-exact selection order for the drop-in `hamming_top_10` /
// `hamming_top_cand` (np_sims/ufunc.c:108-195, :444-551).
//
// The reference's queue is history dependent: a row enters only when its
// distance is strictly below the current worst, and it overwrites the FIRST slot holding the
// worst.  Which of several rows tied at the k-th distance survive, and the slot order of the
// returned ids, therefore cannot be obtained from a sort.  This kernel reproduces the
// sequence on the device, after K5 has produced the exact distances at full HBM speed:
//
//   * one CTA per query streams the distance row (uint16/uint32) in 32 KB tiles, 16-byte
//     coalesced loads, next tile prefetched in registers;
//   * each thread tests its 8 (or 4) consecutive distances against the worst distance at
//     tile start (the worst only ever decreases, so this is a superset filter) — one
//     __syncthreads_or per tile decides "nothing to do", the steady state;
//   * tiles that do contain candidates hand them, in row order, to warp 0, which re-filters
//     each group of 32 threads against the live worst in parallel, re-tests the survivors one
//     by one, writes the slot, and rescans the k slots for the first maximum with a warp
//     max-reduce over (distance, -slot).
//
// No host replay, no CPU fallback: the whole queue lives in shared memory.
#include "common.cuh"
#include "replay.cuh"

namespace nps {

constexpr int kReplayThreads = 1024;
constexpr int kReplayVecs = 2;     // 16-byte vectors per thread per tile

template <typename DistT>
__global__ void __launch_bounds__(kReplayThreads, 1) replay_queue_kernel(const DistT* __restrict__ dist,
                                                                         unsigned long long N, uint32_t k,
                                                                         uint64_t* __restrict__ out_rows,
                                                                         uint64_t* __restrict__ out_scores,
                                                                         const uint32_t* __restrict__ gate) {
    if (gate != nullptr && *gate == 0u) return;   // gated fallback launch (ref_scan.cu), nothing to redo
    constexpr int EPV = 16 / sizeof(DistT);                 // elements per vector
    constexpr int EPT = EPV * kReplayVecs;                  // elements per thread per tile
    constexpr uint32_t TILE = kReplayThreads * EPT;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* slot_score = reinterpret_cast<uint32_t*>(smem_raw);            // [k]
    uint64_t* slot_row = reinterpret_cast<uint64_t*>(smem_raw + ((k * 4 + 15) & ~15u));  // [k]
    __shared__ uint32_t flags[kReplayThreads];               // per-thread candidate bits of the tile
    __shared__ DistT tile_s[TILE];
    __shared__ uint32_t worst_s, worst_slot_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DistT* d = dist + (size_t)blockIdx.x * N;
    uint64_t* rows_o = out_rows + (size_t)blockIdx.x * k;

    // ---- fill phase (ufunc.c:153-157): the first min(k,N) rows take slots 0.. in order
    const uint32_t filled = N < k ? (uint32_t)N : k;
    for (uint32_t i = tid; i < k; i += kReplayThreads) {
        slot_score[i] = i < filled ? (uint32_t)d[i] : 0xFFFFFFFFu;
        slot_row[i] = i < filled ? (uint64_t)i : kNone;
    }
    __syncthreads();

    // warp 0: first slot holding the maximum (ufunc.c:163-169).  Lane l owns slots l, l + 32, ... and
    // caches the maximum over them, so an insert costs one re-scan of the OWNER's k/32 slots plus a
    // warp max-reduce, not a pass over all k slots (k = 1000: ~150 instead of ~600 cycles).
    uint64_t lane_best = 0;
    auto rescan_lane = [&]() {
        uint64_t best = 0;
        for (uint32_t i = lane; i < k; i += 32) {
            const uint64_t key = ((uint64_t)((volatile uint32_t*)slot_score)[i] << 32) | (0xFFFFFFFFu - i);
            best = key > best ? key : best;
        }
        lane_best = best;
    };
    auto publish_worst = [&]() {
        uint64_t best = lane_best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) {
            worst_s = (uint32_t)(best >> 32);
            worst_slot_s = 0xFFFFFFFFu - (uint32_t)best;
        }
        __syncwarp();
    };

    if (N > k && k > 0) {
        if (warp == 0) {
            rescan_lane();
            publish_worst();
        }
        __syncthreads();

        const unsigned long long first = k;                       // sequential part starts at row k
        const unsigned long long tile0 = (first / TILE) * TILE;   // aligned tile containing row k
        uint4 cur[kReplayVecs], nxt[kReplayVecs];
        auto load_tile = [&](unsigned long long base, uint4 (&v)[kReplayVecs]) {
#pragma unroll
            for (int j = 0; j < kReplayVecs; ++j) {
                const unsigned long long e = base + ((unsigned long long)j * kReplayThreads + tid) * EPV;
                if (e + EPV <= N) {
                    v[j] = __ldg(reinterpret_cast<const uint4*>(d + e));
                } else {
                    DistT tmp[EPV];
#pragma unroll
                    for (int x = 0; x < EPV; ++x) tmp[x] = (e + x < N) ? d[e + x] : (DistT)~(DistT)0;
                    v[j] = *reinterpret_cast<uint4*>(tmp);
                }
            }
        };
        // 16-byte loads need d + e aligned: the distance row of query q starts at q*N elements
        const bool aligned = ((reinterpret_cast<uintptr_t>(d) & 15u) == 0);
        auto load_tile_safe = [&](unsigned long long base, uint4 (&v)[kReplayVecs]) {
            if (aligned) { load_tile(base, v); return; }
#pragma unroll
            for (int j = 0; j < kReplayVecs; ++j) {
                const unsigned long long e = base + ((unsigned long long)j * kReplayThreads + tid) * EPV;
                DistT tmp[EPV];
#pragma unroll
                for (int x = 0; x < EPV; ++x) tmp[x] = (e + x < N) ? d[e + x] : (DistT)~(DistT)0;
                v[j] = *reinterpret_cast<uint4*>(tmp);
            }
        };

        load_tile_safe(tile0, cur);
        for (unsigned long long base = tile0; base < N; base += TILE) {
            if (base + TILE < N) load_tile_safe(base + TILE, nxt);
            const uint32_t worst = worst_s;
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < kReplayVecs; ++j) {
                const DistT* e = reinterpret_cast<const DistT*>(&cur[j]);
                const unsigned long long e0 = base + ((unsigned long long)j * kReplayThreads + tid) * EPV;
#pragma unroll
                for (int x = 0; x < EPV; ++x)
                    if ((uint32_t)e[x] < worst && e0 + x >= first && e0 + x < N) bits |= 1u << (j * EPV + x);
            }
            if (__syncthreads_or(bits != 0)) {
                // slow path: publish the tile and the flags, let warp 0 replay in row order
#pragma unroll
                for (int j = 0; j < kReplayVecs; ++j)
                    *reinterpret_cast<uint4*>(&tile_s[((size_t)j * kReplayThreads + tid) * EPV]) = cur[j];
                flags[tid] = bits;
                __syncthreads();
                if (warp == 0) {
                    // element order inside the tile: vector j (outer), thread, element x
                    for (int j = 0; j < kReplayVecs; ++j) {
                        for (int t0 = 0; t0 < kReplayThreads; t0 += 32) {
                            uint32_t f = (flags[t0 + lane] >> (j * EPV)) & ((1u << EPV) - 1u);
                            if (f) {
                                // Re-filter against the LIVE worst (it only decreases) before walking the group
                                // in row order: the tile-start filter lets half of the first tile through, and
                                // every candidate costs ~250 cycles of serial work on this warp.
                                const uint32_t w = *reinterpret_cast<volatile uint32_t*>(&worst_s);
                                const uint4 v = *reinterpret_cast<const uint4*>(
                                    &tile_s[((size_t)j * kReplayThreads + t0 + lane) * EPV]);
                                const DistT* e = reinterpret_cast<const DistT*>(&v);
                                uint32_t g = 0;
#pragma unroll
                                for (int x = 0; x < EPV; ++x)
                                    if ((uint32_t)e[x] < w) g |= 1u << x;
                                f &= g;
                            }
                            uint32_t any = __ballot_sync(0xffffffffu, f != 0);
                            while (any) {
                                const int src = __ffs(any) - 1;
                                any &= any - 1;
                                uint32_t fm = __shfl_sync(0xffffffffu, f, src);
                                while (fm) {
                                    const int x = __ffs(fm) - 1;
                                    fm &= fm - 1;
                                    const uint32_t off = ((uint32_t)j * kReplayThreads + t0 + src) * EPV + x;
                                    const uint32_t score = (uint32_t)tile_s[off];
                                    if (score < *reinterpret_cast<volatile uint32_t*>(&worst_s)) {   // ufunc.c:152, live worst
                                        const uint32_t slot = *reinterpret_cast<volatile uint32_t*>(&worst_slot_s);
                                        __syncwarp();
                                        if (lane == 0) {
                                            slot_score[slot] = score;
                                            slot_row[slot] = base + off;
                                        }
                                        __syncwarp();
                                        if ((slot & 31u) == (uint32_t)lane) rescan_lane();
                                        publish_worst();
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int j = 0; j < kReplayVecs; ++j) cur[j] = nxt[j];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < k; i += kReplayThreads) {
        rows_o[i] = slot_row[i];
        if (out_scores) out_scores[(size_t)blockIdx.x * k + i] = slot_row[i] == kNone ? kNone : (uint64_t)slot_score[i];
    }
}

template <typename DistT>
static cudaError_t launch_t(const void* dist, unsigned long long N, uint32_t nq, uint32_t k, uint64_t* out_rows,
                            uint64_t* out_scores, cudaStream_t stream, const uint32_t* gate) {
    const size_t smem = ((k * 4 + 15) & ~15u) + (size_t)k * 8;
    auto kern = replay_queue_kernel<DistT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<nq, kReplayThreads, smem, stream>>>((const DistT*)dist, N, k, out_rows, out_scores, gate);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_replay_queue(const void* dist, int dist_bytes, unsigned long long N, uint32_t nq, uint32_t k,
                                uint64_t* out_rows, uint64_t* out_scores, cudaStream_t stream, const uint32_t* gate) {
    if (nq == 0 || k == 0) return cudaSuccess;
    if (dist_bytes == 2) return launch_t<uint16_t>(dist, N, nq, k, out_rows, out_scores, stream, gate);
    if (dist_bytes == 4) return launch_t<uint32_t>(dist, N, nq, k, out_rows, out_scores, stream, gate);
    return cudaErrorInvalidValue;
}

}  // namespace nps
