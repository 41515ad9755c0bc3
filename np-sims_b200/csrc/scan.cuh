// scan.cuh — K2: fused Hamming scan + per-CTA exact top-k  (replaces np_sims/ufunc.c:150-551).
//
// Data layout in HBM: signatures uint64[N, W] row-major, exactly the reference's table
// (np_sims/lsh.py:39).  Rows are contiguous, so a tile of TILE_ROWS rows is ONE contiguous
// byte range and is fetched with a single 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) into
// a shared-memory ring; a producer warp runs ahead of 8 consumer warps through full/empty
// mbarriers.  Each consumer thread owns R rows of the tile in registers and sweeps the
// query tile (staged once per work item in shared memory, read as 128-bit broadcasts):
// XOR + POPC per 32-bit half word, one IMAD to form the 32-bit selection key
//       key = distance << row_bits | row_in_chunk
// so that integer '<' on keys IS the (distance, lowest index) order of the canonical
// contract.  A key enters the per-(CTA,query) candidate list only if it beats the list's
// current k-th key (tau); that test is one compare + one warp vote per (thread, query) and
// almost always fails, so selection costs nothing in steady state.  Winners are appended
// under a per-query shared-memory lock by the whole (converged) warp; when the list is full
// it is bitonic-sorted by that warp and cut back to k, which tightens tau.
//
// Work decomposition: work item = (row chunk, query tile); items are dealt round-robin to a
// persistent grid of one CTA per SM.  Every item writes its k best keys (as 64-bit
// dist<<40|global_row, ascending) to the workspace; K3 (merge.cu) reduces the chunks.
#pragma once
#include "common.cuh"

namespace nps {

struct ScanParams {
    const uint64_t* hashes;    // [N, W]
    const uint64_t* queries;   // [Q, W]
    uint64_t* out_keys;        // [Q, num_chunks, k]
    const uint32_t* gate;      // non-null: the launch is a no-op unless *gate != 0 (overflow repair)
    unsigned long long N;
    unsigned long long row_base;  // added to every emitted row id (shard offset)
    uint32_t Q, k, cap, words;
    uint32_t chunk_rows;       // rows per chunk: multiple of the tile rows, <= 2^row_bits
    uint32_t num_chunks, q_tile, num_qtiles;
    uint32_t row_bits;         // key = dist << row_bits | row_in_chunk
    uint32_t stages;
    uint32_t tile_rows;        // generic-width kernel only
    uint32_t lists_per_q;      // 1: one shared (locked) candidate list per query; kConsumerWarps: one per warp
};

// Shared-memory carve-up, shared by host (sizing) and device.
struct ScanSmem {
    uint32_t stage_off, q_off, list_off, tau_off, cnt_off, lock_off, bar_off, total;
};
__host__ __device__ inline uint32_t scan_q_stride(uint32_t words) { return ((words + 1) / 2 + 8) * 16; }
__host__ __device__ inline ScanSmem scan_smem_layout(uint32_t tile_bytes, uint32_t stages, uint32_t words,
                                                     uint32_t q_tile, uint32_t cap, uint32_t lists_per_q = 1) {
    ScanSmem s;
    uint32_t off = 0;
    s.stage_off = off;
    off += stages * ((tile_bytes + 127u) & ~127u);
    s.q_off = off;
    off += q_tile * scan_q_stride(words);
    s.list_off = off;
    off += q_tile * lists_per_q * cap * 4;
    s.tau_off = off;
    off += q_tile * lists_per_q * 4;
    s.cnt_off = off;
    off += q_tile * lists_per_q * 4;
    s.lock_off = off;
    off += q_tile * lists_per_q * 4;
    off = (off + 7u) & ~7u;
    s.bar_off = off;
    off += stages * 16;
    s.total = off;
    return s;
}

// ---- warp-level bitonic sort of `cap` (power of two, >= 64) keys in shared memory ----------
__device__ __forceinline__ void warp_sort_list(uint32_t* list, uint32_t cap, int lane) {
    volatile uint32_t* v = list;
    for (uint32_t size = 2; size <= cap; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = lane; i < (cap >> 1); i += 32) {
                const uint32_t pos = 2 * i - (i & (stride - 1));
                const uint32_t a = v[pos], b = v[pos + stride];
                const bool up = (pos & size) == 0;
                if ((a > b) == up) {
                    v[pos] = b;
                    v[pos + stride] = a;
                }
            }
            __syncwarp();
        }
    }
}

// Sort the list, keep the k best.  Returns the new element count; *tau_out = k-th key if full.
// Only the smallest power of two >= cnt (>= 64) is sorted: the early refreshes of tau (below) handle
// short lists.
__device__ __forceinline__ uint32_t warp_compact(uint32_t* list, uint32_t cnt, uint32_t k, uint32_t cap, int lane,
                                                 uint32_t* tau_out) {
    uint32_t n = 64;
    while (n < cnt) n <<= 1;
    if (n > cap) n = cap;
    for (uint32_t i = cnt + lane; i < n; i += 32) list[i] = kEmptyKey;
    __syncwarp();
    warp_sort_list(list, n, lane);
    if (cnt >= k) {
        *tau_out = ((volatile uint32_t*)list)[k - 1];
        return k;
    }
    return cnt;
}

// Append the warp's candidates for query `q` (called with all 32 lanes converged).
template <int R>
__device__ __noinline__ void append_candidates(uint32_t* list, uint32_t* tau_p, uint32_t* cnt_p, uint32_t* lock_p,
                                               const uint32_t (&key)[R], uint32_t valid_mask, uint32_t k,
                                               uint32_t cap, int lane) {
    if (lane == 0) {
        while (atomicCAS(lock_p, 0u, 1u) != 0u) {
        }
    }
    __syncwarp();
    __threadfence_block();
    uint32_t tau = lds32_volatile(tau_p);
    uint32_t cnt = lds32_volatile(cnt_p);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        bool pred = ((valid_mask >> r) & 1u) && key[r] < tau;
        uint32_t mask = __ballot_sync(0xffffffffu, pred);
        if (mask == 0) continue;
        uint32_t n = __popc(mask);
        if (cnt + n > cap) {
            cnt = warp_compact(list, cnt, k, cap, lane, &tau);
            pred = pred && key[r] < tau;
            mask = __ballot_sync(0xffffffffu, pred);
            n = __popc(mask);
        }
        if (pred) ((volatile uint32_t*)list)[cnt + __popc(mask & ((1u << lane) - 1u))] = key[r];
        cnt += n;
        __syncwarp();
    }
    // Tighten tau as soon as k candidates exist, and again whenever k more have arrived: tau is what
    // every warp filters with, and a threshold that is refreshed only when the list is full (cap - k
    // appends later) stays at the k-th best of the first few thousand rows for the rest of the chunk —
    // measured on a single-query scan: ~5 lock-serialised appends per tile, 0.3 ms for a 16 M-row table
    // whatever the signature width (128 MB .. 1.3 GB), i.e. bound by this path and not by HBM.
    if (cnt >= k && (tau == kEmptyKey || cnt >= 2u * k)) cnt = warp_compact(list, cnt, k, cap, lane, &tau);
    __syncwarp();
    if (lane == 0) {
        sts32_volatile(cnt_p, cnt);
        sts32_volatile(tau_p, tau);
        __threadfence_block();
        atomicExch(lock_p, 0u);
    }
    __syncwarp();
}

// Producer: one lane streams the tiles of every work item of this CTA through the ring.
template <uint32_t ROW_BYTES>
__device__ __forceinline__ void scan_producer(const ScanParams& p, uint32_t tile_rows, uint8_t* stage_base,
                                              uint32_t stage_stride, uint64_t* full, uint64_t* empty,
                                              uint32_t row_bytes_rt) {
    const uint32_t row_bytes = ROW_BYTES ? ROW_BYTES : row_bytes_rt;
    const uint32_t total_items = p.num_chunks * p.num_qtiles;
    const uint32_t S = p.stages;
    uint32_t it = 0, s = 0, ph = 0;
    for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const uint32_t chunk = item / p.num_qtiles;
        const unsigned long long row0 = (unsigned long long)chunk * p.chunk_rows;
        unsigned long long row1 = row0 + p.chunk_rows;
        if (row1 > p.N) row1 = p.N;
        for (unsigned long long r = row0; r < row1; r += tile_rows, ++it) {
            mbar_wait(&empty[s], ph ^ 1u);
            const unsigned long long left = row1 - r;
            const uint32_t rows = left < tile_rows ? (uint32_t)left : tile_rows;
            const uint32_t bytes = rows * row_bytes;
            const uint32_t bulk = bytes & ~15u;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.hashes) + r * row_bytes;
            uint8_t* dst = stage_base + s * stage_stride;
            if (bytes != bulk)  // odd row count x odd width: last 8 bytes by hand
                *reinterpret_cast<volatile uint64_t*>(dst + bulk) = *reinterpret_cast<const uint64_t*>(src + bulk);
            mbar_arrive_expect_tx(&full[s], bulk);
            if (bulk) bulk_g2s(dst, src, bulk, &full[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }
    }
}

// =============================================================================================
// Width-specialised kernel: W 64-bit words per signature, R rows per thread in registers.
// =============================================================================================
template <int W, int R>
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int TILE_ROWS = kConsumerThreads * R;
    constexpr uint32_t ROW_BYTES = W * 8;
    constexpr uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;
    constexpr bool EVEN = (W % 2) == 0;
    constexpr int V = W / 2;                       // 16-byte vectors per row (even W)
    constexpr int VQ = (W + 1) / 2;                // 16-byte vectors per query (zero padded)
    constexpr int G = EVEN ? gcd_int(V, 8) : 1;    // bank-conflict rotation period
    const uint32_t QSTRIDE = scan_q_stride(W);

    if (p.gate != nullptr && *p.gate == 0u) return;   // gated repair launch, nothing to repair
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t LPQ = p.lists_per_q;   // > 1: every consumer warp owns a private list per query (no lock contention)
    const ScanSmem L = scan_smem_layout(TILE_BYTES, p.stages, W, p.q_tile, p.cap, LPQ);
    uint8_t* stage_base = smem + L.stage_off;
    const uint32_t stage_stride = (TILE_BYTES + 127u) & ~127u;
    uint8_t* q_s = smem + L.q_off;
    uint32_t* list_s = reinterpret_cast<uint32_t*>(smem + L.list_off);
    uint32_t* tau_s = reinterpret_cast<uint32_t*>(smem + L.tau_off);
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem + L.cnt_off);
    uint32_t* lock_s = reinterpret_cast<uint32_t*>(smem + L.lock_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0) scan_producer<ROW_BYTES>(p, TILE_ROWS, stage_base, stage_stride, full, empty, 0);
        return;
    }

    const uint32_t total_items = p.num_chunks * p.num_qtiles;
    const uint32_t S = p.stages, k = p.k, cap = p.cap;
    const uint32_t key_mul = 1u << p.row_bits;
    const int rot = EVEN ? (((tid & 7) * G) >> 3) : 0;
    const uint32_t q_base = smem_u32(q_s) + rot * 16;
    uint32_t s = 0, ph = 0;

    for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const uint32_t chunk = item / p.num_qtiles, qt = item - chunk * p.num_qtiles;
        const uint32_t q0 = qt * p.q_tile;
        const uint32_t nq = (p.Q - q0) < p.q_tile ? (p.Q - q0) : p.q_tile;
        const unsigned long long row0 = (unsigned long long)chunk * p.chunk_rows;
        unsigned long long row1 = row0 + p.chunk_rows;
        if (row1 > p.N) row1 = p.N;

        // ---- stage the query tile: each query as VQ 16-byte vectors + 8 wrap-around copies
        consumer_bar();  // previous item's lists fully written out
        for (uint32_t i = tid; i < nq * (VQ + 8); i += kConsumerThreads) {
            const uint32_t q = i / (VQ + 8), v = i - q * (VQ + 8);
            const uint32_t vv = v < VQ ? v : (v - VQ) % VQ;
            const uint64_t* src = p.queries + (unsigned long long)(q0 + q) * W;
            ulonglong2 val;
            val.x = src[2 * vv];
            val.y = (2 * vv + 1 < W) ? src[2 * vv + 1] : 0ull;
            *reinterpret_cast<ulonglong2*>(q_s + q * QSTRIDE + v * 16) = val;
        }
        for (uint32_t q = tid; q < nq * LPQ; q += kConsumerThreads) {
            tau_s[q] = kEmptyKey;
            cnt_s[q] = 0;
            lock_s[q] = 0;
        }
        consumer_bar();
        const uint32_t my_list = LPQ > 1 ? (uint32_t)warp : 0u;   // list (q, my_list) of query q

        uint32_t tile_row = 0;  // row offset of the current tile inside the chunk
        for (unsigned long long r0 = row0; r0 < row1; r0 += TILE_ROWS, tile_row += TILE_ROWS) {
            mbar_wait(&full[s], ph);
            const uint32_t st = smem_u32(stage_base + s * stage_stride);

            // ---- my R rows: shared -> registers (conflict-free: see DESIGN.md "bank rotation")
            uint32_t h[R][2 * W];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t row = r * kConsumerThreads + tid;
                if constexpr (EVEN) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        int jj = v + rot;
                        if (jj >= V) jj -= V;
                        const uint4 x = lds128(st + (row * V + jj) * 16);
                        h[r][4 * v + 0] = x.x;
                        h[r][4 * v + 1] = x.y;
                        h[r][4 * v + 2] = x.z;
                        h[r][4 * v + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const uint2 x = lds64(st + (row * W + w) * 8);
                        h[r][2 * w + 0] = x.x;
                        h[r][2 * w + 1] = x.y;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }

            uint32_t lrow[R];
            uint32_t valid_mask = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                lrow[r] = tile_row + r * kConsumerThreads + tid;
                if (row0 + lrow[r] < row1) valid_mask |= 1u << r;
            }

            // ---- sweep the query tile
            uint32_t qaddr = q_base;
            for (uint32_t q = 0; q < nq; ++q, qaddr += QSTRIDE) {
                const uint32_t li = q * LPQ + my_list;
                const uint32_t tau = lds32_volatile(&tau_s[li]);
                uint32_t d[R];
#pragma unroll
                for (int r = 0; r < R; ++r) d[r] = 0;
#pragma unroll
                for (int v = 0; v < VQ; ++v) {
                    const uint4 qv = lds128(qaddr + v * 16);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        d[r] += __popc(h[r][4 * v + 0] ^ qv.x) + __popc(h[r][4 * v + 1] ^ qv.y);
                        if (4 * v + 2 < 2 * W)
                            d[r] += __popc(h[r][4 * v + 2] ^ qv.z) + __popc(h[r][4 * v + 3] ^ qv.w);
                    }
                }
                uint32_t key[R];
                bool hit = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    key[r] = d[r] * key_mul + lrow[r];
                    hit |= key[r] < tau;
                }
                if (__any_sync(0xffffffffu, hit))
                    append_candidates<R>(list_s + li * cap, &tau_s[li], &cnt_s[li], &lock_s[li], key, valid_mask, k, cap,
                                         lane);
            }
        }

        // ---- item done: sort every list, emit k keys per query as dist<<40 | global row
        consumer_bar();
        if (LPQ > 1) {
            // private lists: every warp sorts its own, then one warp per query merges the LPQ k-prefixes
            // (LPQ * k <= 256 keys) in the query's block of lists and leaves the result in list (q, 0)
            for (uint32_t q = 0; q < nq; ++q) {
                const uint32_t li = q * LPQ + (uint32_t)warp;
                uint32_t tau_unused;
                const uint32_t c = warp_compact(list_s + li * cap, cnt_s[li], k, cap, lane, &tau_unused);
                if (lane == 0) cnt_s[li] = c;
            }
            consumer_bar();
            for (uint32_t q = warp; q < nq; q += kConsumerWarps) {
                uint32_t* block = list_s + q * LPQ * cap;
                const uint32_t total = LPQ * k;                 // <= 256
                uint32_t vals[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t e = (uint32_t)j * 32u + lane;
                    const uint32_t src = e / k, idx = e - src * k;
                    vals[j] = (e < total && idx < cnt_s[q * LPQ + src]) ? block[src * cap + idx] : kEmptyKey;
                }
                __syncwarp();
                uint32_t n = 64;
                while (n < total) n <<= 1;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t e = (uint32_t)j * 32u + lane;
                    if (e < n) block[e] = vals[j];
                }
                __syncwarp();
                warp_sort_list(block, n, lane);
                uint32_t c = 0;
                for (uint32_t i = lane; i < k; i += 32) c += block[i] != kEmptyKey;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if (lane == 0) cnt_s[q * LPQ] = c;
                __syncwarp();
            }
        }
        for (uint32_t q = warp; q < nq; q += kConsumerWarps) {
            uint32_t* list = list_s + q * LPQ * cap;
            uint32_t tau_unused;
            const uint32_t cnt = LPQ > 1 ? cnt_s[q * LPQ] : warp_compact(list, cnt_s[q], k, cap, lane, &tau_unused);
            uint64_t* out = p.out_keys + ((unsigned long long)(q0 + q) * p.num_chunks + chunk) * k;
            for (uint32_t i = lane; i < k; i += 32) {
                uint64_t o = kNone;
                if (i < cnt) {
                    const uint32_t key = list[i];
                    const uint64_t dist = key >> p.row_bits;
                    const uint64_t row = p.row_base + row0 + (key & (key_mul - 1u));
                    o = (dist << kGlobalRowBits) | row;
                }
                out[i] = o;
            }
        }
    }
}

#ifdef NPS_SCAN_GENERIC
// =============================================================================================
// Generic-width kernel (W > 32, e.g. the reference's 2560-bit signatures): rows stay in shared
// memory, one row per thread, runtime word loop.  Same selection machinery.
// =============================================================================================
__global__ void __launch_bounds__(kScanThreads, 1) scan_topk_generic_kernel(const ScanParams p) {
    if (p.gate != nullptr && *p.gate == 0u) return;   // gated repair launch, nothing to repair
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t W = p.words, ROW_BYTES = W * 8, TILE_ROWS = p.tile_rows;
    const uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;
    const uint32_t QSTRIDE = scan_q_stride(W);
    const ScanSmem L = scan_smem_layout(TILE_BYTES, p.stages, W, p.q_tile, p.cap);
    uint8_t* stage_base = smem + L.stage_off;
    const uint32_t stage_stride = (TILE_BYTES + 127u) & ~127u;
    uint8_t* q_s = smem + L.q_off;
    uint32_t* list_s = reinterpret_cast<uint32_t*>(smem + L.list_off);
    uint32_t* tau_s = reinterpret_cast<uint32_t*>(smem + L.tau_off);
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(smem + L.cnt_off);
    uint32_t* lock_s = reinterpret_cast<uint32_t*>(smem + L.lock_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == kConsumerWarps) {
        if (lane == 0) scan_producer<0>(p, TILE_ROWS, stage_base, stage_stride, full, empty, ROW_BYTES);
        return;
    }

    const uint32_t total_items = p.num_chunks * p.num_qtiles;
    const uint32_t S = p.stages, k = p.k, cap = p.cap;
    const uint32_t key_mul = 1u << p.row_bits;
    uint32_t s = 0, ph = 0;
    for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const uint32_t chunk = item / p.num_qtiles, qt = item - chunk * p.num_qtiles;
        const uint32_t q0 = qt * p.q_tile;
        const uint32_t nq = (p.Q - q0) < p.q_tile ? (p.Q - q0) : p.q_tile;
        const unsigned long long row0 = (unsigned long long)chunk * p.chunk_rows;
        unsigned long long row1 = row0 + p.chunk_rows;
        if (row1 > p.N) row1 = p.N;

        consumer_bar();
        for (uint32_t i = tid; i < nq * W; i += kConsumerThreads) {
            const uint32_t q = i / W, w = i - q * W;
            *reinterpret_cast<uint64_t*>(q_s + q * QSTRIDE + w * 8) = p.queries[(unsigned long long)(q0 + q) * W + w];
        }
        for (uint32_t q = tid; q < nq; q += kConsumerThreads) {
            tau_s[q] = kEmptyKey;
            cnt_s[q] = 0;
            lock_s[q] = 0;
        }
        consumer_bar();

        uint32_t tile_row = 0;
        for (unsigned long long r0 = row0; r0 < row1; r0 += TILE_ROWS, tile_row += TILE_ROWS) {
            mbar_wait(&full[s], ph);
            const uint64_t* rows = reinterpret_cast<const uint64_t*>(stage_base + s * stage_stride);
            const uint32_t lrow = tile_row + tid;
            const bool valid = (uint32_t)tid < TILE_ROWS && row0 + lrow < row1;
            const uint64_t* my = rows + (size_t)((uint32_t)tid < TILE_ROWS ? tid : 0) * W;
            for (uint32_t q = 0; q < nq; ++q) {
                const uint32_t tau = lds32_volatile(&tau_s[q]);
                const uint64_t* qq = reinterpret_cast<const uint64_t*>(q_s + q * QSTRIDE);
                uint32_t d = 0;
                for (uint32_t w = 0; w < W; ++w) d += __popcll(my[w] ^ qq[w]);
                uint32_t key[1] = {d * key_mul + lrow};
                const bool hit = valid && key[0] < tau;
                if (__any_sync(0xffffffffu, hit))
                    append_candidates<1>(list_s + q * cap, &tau_s[q], &cnt_s[q], &lock_s[q], key, valid ? 1u : 0u, k,
                                         cap, lane);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }

        consumer_bar();
        for (uint32_t q = warp; q < nq; q += kConsumerWarps) {
            uint32_t* list = list_s + q * cap;
            uint32_t tau_unused;
            const uint32_t cnt = warp_compact(list, cnt_s[q], k, cap, lane, &tau_unused);
            uint64_t* out = p.out_keys + ((unsigned long long)(q0 + q) * p.num_chunks + chunk) * k;
            for (uint32_t i = lane; i < k; i += 32) {
                uint64_t o = kNone;
                if (i < cnt) {
                    const uint32_t key = list[i];
                    const uint64_t dist = key >> p.row_bits;
                    const uint64_t row = p.row_base + row0 + (key & (key_mul - 1u));
                    o = (dist << kGlobalRowBits) | row;
                }
                out[i] = o;
            }
        }
    }
}

#endif  // NPS_SCAN_GENERIC

// host-side launcher table (scan_inst.cu)
typedef cudaError_t (*ScanLaunchFn)(const ScanParams&, int grid, size_t smem, cudaStream_t);
ScanLaunchFn scan_launcher_for_width(int words);   // nullptr if no specialisation
cudaError_t launch_scan_generic(const ScanParams&, int grid, size_t smem, cudaStream_t);

}  // namespace nps
