// hist_select.cuh — K2H: canonical top-k for FEW queries and LARGE k (k > 32) at HBM speed.
// Replaces, for that regime, the sorted shared-memory candidate lists of K2 (scan.cuh), whose lock and
// re-sorts made k = 100 / 1000 scans 4x / 15x slower than the table stream (profiles/r01_c5_sweep.txt).
// Reference semantics: the k rows with the smallest Hamming distance (np_sims/ufunc.c:174-195 keeps
// exactly those; here ties are broken by lowest row id — the canonical contract).
//
// A Hamming distance over W <= 63 words is an integer < 4096, so selection is a counting problem:
//
//   H1  hist_scan_kernel   (persistent, one CTA per SM; same TMA ring + register-tile XOR+POPC sweep as
//       K2): streams the table ONCE; per (row, query) it stores the distance as uint16 into the workspace
//       and bumps a shared-memory histogram, flushed to the query's global histogram per work item.
//       Algorithmic bytes: N*W*8 read + Q*N*2 written.
//   H2  hist_count_kernel  : every CTA derives the exact k-th smallest distance d_k and the number of rows
//       below it from the histogram, then counts, per warp-sized row range, the rows with d < d_k and
//       d == d_k (reads the uint16 distances: Q*N*2 bytes, mostly from L2).
//   H3  hist_emit_kernel   : exclusive prefix over those counts -> every warp knows where its rows go;
//       rows with d < d_k are all kept, of the rows with d == d_k the first (k - #less) IN ROW ORDER.
//       The last CTA of a query sorts its k keys (distance << 40 | row) and writes ids / distances.
//
// Exact for every input (heavy ties, fewer than k rows, duplicate rows); no candidate buffers that can
// overflow.  Distances are recomputed nowhere: H2/H3 work on 2 bytes per row.
#pragma once
#include "common.cuh"

namespace nps {

constexpr uint32_t kHistMaxWords = 63;
constexpr int kHistSelThreads = 256;
constexpr uint32_t kHistWarpRows = 4096;                       // rows per warp range in H2 / H3
constexpr uint32_t kHistCtaRows = kHistWarpRows * (kHistSelThreads / 32);

struct HistParams {
    const uint64_t* hashes;    // [N, W]
    const uint64_t* queries;   // [Q, W]
    uint16_t* dist;            // [Q, n_pad] distances (n_pad = N rounded up to 8)
    uint32_t* ghist;           // [Q, bins] zero on entry
    unsigned long long N, n_pad;
    uint32_t Q, words, bins;   // bins = 64 * W + 1 rounded up to a multiple of 32
    uint32_t chunk_rows, num_chunks, q_tile, num_qtiles, stages, tile_rows;
};

struct HistSmem {
    uint32_t stage_off, q_off, hist_off, bar_off, total;
};
__host__ __device__ inline uint32_t hist_q_stride(uint32_t words) { return ((words + 1) / 2 + 8) * 16; }
__host__ __device__ inline HistSmem hist_smem_layout(uint32_t tile_bytes, uint32_t stages, uint32_t words, uint32_t q_tile,
                                                     uint32_t bins) {
    HistSmem s;
    uint32_t off = 0;
    s.stage_off = off;
    off += stages * ((tile_bytes + 127u) & ~127u);
    s.q_off = off;
    off += q_tile * hist_q_stride(words);
    s.hist_off = off;
    off += q_tile * bins * 4;
    off = (off + 7u) & ~7u;
    s.bar_off = off;
    off += stages * 16;
    s.total = off;
    return s;
}

#ifdef NPS_HIST_DEVICE_CODE
__device__ __forceinline__ void hist_producer(const HistParams& p, uint32_t tile_rows, uint32_t row_bytes, uint8_t* stage_base,
                                              uint32_t stage_stride, uint64_t* full, uint64_t* empty) {
    const uint32_t total_items = p.num_chunks * p.num_qtiles;
    const uint32_t S = p.stages;
    uint32_t s = 0, ph = 0;
    for (uint32_t item = blockIdx.x; item < total_items; item += gridDim.x) {
        const uint32_t chunk = item / p.num_qtiles;
        const unsigned long long row0 = (unsigned long long)chunk * p.chunk_rows;
        unsigned long long row1 = row0 + p.chunk_rows;
        if (row1 > p.N) row1 = p.N;
        for (unsigned long long r = row0; r < row1; r += tile_rows) {
            mbar_wait(&empty[s], ph ^ 1u);
            const unsigned long long left = row1 - r;
            const uint32_t rows = left < tile_rows ? (uint32_t)left : tile_rows;
            const uint32_t bytes = rows * row_bytes, bulk = bytes & ~15u;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.hashes) + r * row_bytes;
            uint8_t* dst = stage_base + s * stage_stride;
            if (bytes != bulk)
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

// shared-memory histograms of the item -> the queries' global histograms; leaves them zeroed
__device__ __forceinline__ void hist_flush(uint32_t* hist_s, uint32_t* ghist, uint32_t q0, uint32_t nq, uint32_t bins, int tid) {
    consumer_bar();
    for (uint32_t i = tid; i < nq * bins; i += kConsumerThreads) {
        const uint32_t c = hist_s[i];
        if (c) {
            atomicAdd(&ghist[(size_t)q0 * bins + i], c);   // (q0 + i / bins) * bins + i % bins
            hist_s[i] = 0;
        }
    }
}

// =============================================================================================
// H1, width-specialised: W words per signature, R rows per thread in registers (as K2, scan.cuh)
// =============================================================================================
template <int W, int R>
__global__ void __launch_bounds__(kScanThreads, 1) hist_scan_kernel(const HistParams p) {
    constexpr int TILE_ROWS = kConsumerThreads * R;
    constexpr uint32_t ROW_BYTES = W * 8;
    constexpr uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;
    constexpr bool EVEN = (W % 2) == 0;
    constexpr int V = W / 2;
    constexpr int VQ = (W + 1) / 2;
    constexpr int G = EVEN ? gcd_int(V, 8) : 1;
    const uint32_t QSTRIDE = hist_q_stride(W);

    extern __shared__ __align__(128) uint8_t smem[];
    const HistSmem L = hist_smem_layout(TILE_BYTES, p.stages, W, p.q_tile, p.bins);
    uint8_t* stage_base = smem + L.stage_off;
    const uint32_t stage_stride = (TILE_BYTES + 127u) & ~127u;
    uint8_t* q_s = smem + L.q_off;
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(smem + L.hist_off);
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
    for (uint32_t i = tid; i < p.q_tile * p.bins; i += kScanThreads) hist_s[i] = 0;
    __syncthreads();
    if (warp == kConsumerWarps) {
        if (lane == 0) hist_producer(p, TILE_ROWS, ROW_BYTES, stage_base, stage_stride, full, empty);
        return;
    }
    const uint32_t total_items = p.num_chunks * p.num_qtiles;
    const uint32_t S = p.stages, bins = p.bins;
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
        consumer_bar();   // previous item's histogram flushed, its query tile no longer read
        for (uint32_t i = tid; i < nq * (VQ + 8); i += kConsumerThreads) {
            const uint32_t q = i / (VQ + 8), v = i - q * (VQ + 8);
            const uint32_t vv = v < VQ ? v : (v - VQ) % VQ;
            const uint64_t* src = p.queries + (unsigned long long)(q0 + q) * W;
            ulonglong2 val;
            val.x = src[2 * vv];
            val.y = (2 * vv + 1 < W) ? src[2 * vv + 1] : 0ull;
            *reinterpret_cast<ulonglong2*>(q_s + q * QSTRIDE + v * 16) = val;
        }
        consumer_bar();
        for (unsigned long long r0 = row0; r0 < row1; r0 += TILE_ROWS) {
            mbar_wait(&full[s], ph);
            const uint32_t st = smem_u32(stage_base + s * stage_stride);
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
            uint32_t qaddr = q_base;
            for (uint32_t q = 0; q < nq; ++q, qaddr += QSTRIDE) {
                uint32_t d[R];
#pragma unroll
                for (int r = 0; r < R; ++r) d[r] = 0;
#pragma unroll
                for (int v = 0; v < VQ; ++v) {
                    const uint4 qv = lds128(qaddr + v * 16);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        d[r] += __popc(h[r][4 * v + 0] ^ qv.x) + __popc(h[r][4 * v + 1] ^ qv.y);
                        if (4 * v + 2 < 2 * W) d[r] += __popc(h[r][4 * v + 2] ^ qv.z) + __popc(h[r][4 * v + 3] ^ qv.w);
                    }
                }
                uint16_t* drow = p.dist + (size_t)(q0 + q) * p.n_pad;
                uint32_t* hq = hist_s + q * bins;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned long long row = r0 + (uint32_t)(r * kConsumerThreads + tid);
                    if (row < row1) {
                        drow[row] = (uint16_t)d[r];
                        atomicAdd(&hq[d[r]], 1u);
                    }
                }
            }
        }
        hist_flush(hist_s, p.ghist, q0, nq, bins, tid);
    }
}

#ifdef NPS_SCAN_GENERIC
// H1, generic width (32 < W <= 63): rows stay in shared memory, one row per thread.
__global__ void __launch_bounds__(kScanThreads, 1) hist_scan_generic_kernel(const HistParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t W = p.words, ROW_BYTES = W * 8, TILE_ROWS = p.tile_rows;
    const uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;
    const uint32_t QSTRIDE = hist_q_stride(W);
    const HistSmem L = hist_smem_layout(TILE_BYTES, p.stages, W, p.q_tile, p.bins);
    uint8_t* stage_base = smem + L.stage_off;
    const uint32_t stage_stride = (TILE_BYTES + 127u) & ~127u;
    uint8_t* q_s = smem + L.q_off;
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(smem + L.hist_off);
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
    for (uint32_t i = tid; i < p.q_tile * p.bins; i += kScanThreads) hist_s[i] = 0;
    __syncthreads();
    if (warp == kConsumerWarps) {
        if (lane == 0) hist_producer(p, TILE_ROWS, ROW_BYTES, stage_base, stage_stride, full, empty);
        return;
    }
    const uint32_t total_items = p.num_chunks * p.num_qtiles;
    const uint32_t S = p.stages, bins = p.bins;
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
        consumer_bar();
        for (unsigned long long r0 = row0; r0 < row1; r0 += TILE_ROWS) {
            mbar_wait(&full[s], ph);
            const uint64_t* rows = reinterpret_cast<const uint64_t*>(stage_base + s * stage_stride);
            const unsigned long long row = r0 + (uint32_t)tid;
            const bool valid = (uint32_t)tid < TILE_ROWS && row < row1;
            const uint64_t* my = rows + (size_t)((uint32_t)tid < TILE_ROWS ? tid : 0) * W;
            for (uint32_t q = 0; q < nq; ++q) {
                const uint64_t* qq = reinterpret_cast<const uint64_t*>(q_s + q * QSTRIDE);
                uint32_t d = 0;
                for (uint32_t w = 0; w < W; ++w) d += __popcll(my[w] ^ qq[w]);
                if (valid) {
                    p.dist[(size_t)(q0 + q) * p.n_pad + row] = (uint16_t)d;
                    atomicAdd(&hist_s[q * bins + d], 1u);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }
        hist_flush(hist_s, p.ghist, q0, nq, bins, tid);
    }
}
#endif  // NPS_SCAN_GENERIC
#endif  // NPS_HIST_DEVICE_CODE

// host-side launcher table (scan_inst.cu)
typedef cudaError_t (*HistLaunchFn)(const HistParams&, int grid, size_t smem, cudaStream_t);
HistLaunchFn hist_launcher_for_width(int words);   // nullptr if no specialisation (words > 32)
cudaError_t launch_hist_scan_generic(const HistParams&, int grid, size_t smem, cudaStream_t);

// ---- host: plan + run (hist_select.cu) ----------------------------------------------------------
struct HistPlan {
    bool applicable;
    uint32_t bins, tile_rows, stages, q_tile, num_qtiles, chunk_rows, num_chunks, grid, smem;
    uint32_t sel_ctas;                    // CTAs per query of H2 / H3
    unsigned long long n_pad;
    size_t off_hist, off_thr, off_done, off_cnt, off_stage, off_dist, zero_bytes, total;
};
HistPlan hist_plan(unsigned long long N, uint32_t W, uint32_t Q, uint32_t k, int sm_count, int smem_optin);
// d_rows / d_dists / d_keys: [Q, k], any may be null.  ws: plan.total bytes, 256-byte aligned.
cudaError_t run_hist_top_n(const HistPlan& pl, const uint64_t* d_hashes, unsigned long long N, uint32_t W,
                           const uint64_t* d_queries, uint32_t Q, uint32_t k, unsigned long long row_base,
                           uint64_t* d_rows, uint64_t* d_dists, uint64_t* d_keys, void* ws, cudaStream_t stream,
                           cudaEvent_t* events);   // nullable: [3] recorded before H1, after H1, after H3

}  // namespace nps
