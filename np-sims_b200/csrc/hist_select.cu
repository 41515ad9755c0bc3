// hist_select.cu — K2H host side + the two selection kernels (H2 count, H3 emit + sort).
// See hist_select.cuh for the design.
#define NPS_HIST_DEVICE_CODE 1
#include "hist_select.cuh"

namespace nps {

struct HistSelParams {
    const uint16_t* dist;      // [Q, n_pad]
    const uint32_t* ghist;     // [Q, bins]
    uint32_t* thr;             // [Q, 4]: d_k, rows below d_k, ties to keep, unused
    uint32_t* done;            // [Q] CTAs of H3 finished (zero on entry)
    uint2* cnt;                // [Q, sel_ctas * 8] per warp range: (rows below d_k, rows equal to d_k)
    uint64_t* stage;           // [Q, k] selected keys, unordered
    uint64_t* out_rows;        // [Q, k] nullable
    uint64_t* out_dists;
    uint64_t* out_keys;
    unsigned long long N, n_pad, row_base;
    uint32_t Q, k, bins, sel_ctas, sort_cap;
};

// d_k = k-th smallest distance, below = #rows with d < d_k; if the table has fewer than k rows every row is
// "below" (d_k = bins).  All threads of the CTA return the same values.
__device__ __forceinline__ void hist_threshold(const uint32_t* __restrict__ gh, uint32_t bins, uint32_t k, uint32_t* sm,
                                               uint32_t& dk, uint32_t& below) {
    // sm: [kHistSelThreads / 32 + 2] scratch
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (bins + kHistSelThreads - 1) / kHistSelThreads;   // <= 16 bins per thread
    const uint32_t b0 = tid * per;
    uint32_t local = 0;
    for (uint32_t j = 0; j < per; ++j)
        if (b0 + j < bins) local += __ldcg(gh + b0 + j);
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    if (tid == 0) {
        sm[kHistSelThreads / 32] = bins;   // defaults: fewer than k rows
        sm[kHistSelThreads / 32 + 1] = 0;
    }
    __syncthreads();
    uint32_t before = inc - local;
    uint32_t total = 0;
    for (int w = 0; w < kHistSelThreads / 32; ++w) {
        const uint32_t s = sm[w];
        if (w < warp) before += s;
        total += s;
    }
    if (before < k && before + local >= k) {
        uint32_t c = before;
        for (uint32_t j = 0; j < per; ++j) {
            const uint32_t b = b0 + j < bins ? __ldcg(gh + b0 + j) : 0u;
            if (c < k && c + b >= k) {
                sm[kHistSelThreads / 32] = b0 + j;
                sm[kHistSelThreads / 32 + 1] = c;
            }
            c += b;
        }
    }
    __syncthreads();
    dk = sm[kHistSelThreads / 32];
    below = total < k ? total : sm[kHistSelThreads / 32 + 1];
    __syncthreads();
}

// rows of a warp range: (less, tie) flags of 8 consecutive distances per lane, 256 rows per iteration
__device__ __forceinline__ void hist_flags(const uint16_t* __restrict__ d, unsigned long long row, unsigned long long N,
                                           uint32_t dk, uint32_t& less, uint32_t& tie) {
    less = tie = 0;
    if (row >= N) return;
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(d + row));   // n_pad: reads inside the row
    const uint16_t* e = reinterpret_cast<const uint16_t*>(&v);
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        if (row + x < N) {
            less |= ((uint32_t)e[x] < dk ? 1u : 0u) << x;
            tie |= ((uint32_t)e[x] == dk ? 1u : 0u) << x;
        }
    }
}

// H2: per warp range (kHistWarpRows rows) the number of rows below / at the k-th distance
__global__ void __launch_bounds__(kHistSelThreads) hist_count_kernel(const HistSelParams p) {
    __shared__ uint32_t sm[kHistSelThreads / 32 + 2];
    const uint32_t q = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t dk, below;
    hist_threshold(p.ghist + (size_t)q * p.bins, p.bins, p.k, sm, dk, below);
    if (blockIdx.x == 0) {
        if (tid == 0) {
            p.thr[4 * q + 0] = dk;
            p.thr[4 * q + 1] = below;
            p.thr[4 * q + 2] = p.k - (below < p.k ? below : p.k);
        }
        for (uint32_t i = tid; i < p.k; i += kHistSelThreads) p.stage[(size_t)q * p.k + i] = kNone;
    }
    const uint16_t* d = p.dist + (size_t)q * p.n_pad;
    const unsigned long long r0 = (unsigned long long)blockIdx.x * kHistCtaRows + (unsigned long long)warp * kHistWarpRows;
    uint32_t nl = 0, nt = 0;
    for (uint32_t i = 0; i < kHistWarpRows; i += 256) {
        uint32_t less, tie;
        hist_flags(d, r0 + i + lane * 8u, p.N, dk, less, tie);
        nl += __popc(less);
        nt += __popc(tie);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
        nt += __shfl_xor_sync(0xffffffffu, nt, o);
    }
    if (lane == 0) p.cnt[((size_t)q * p.sel_ctas + blockIdx.x) * 8 + warp] = make_uint2(nl, nt);
}

// H3: place the selected rows, then the last CTA of the query sorts the k keys
__global__ void __launch_bounds__(kHistSelThreads) hist_emit_kernel(const HistSelParams p) {
    extern __shared__ __align__(16) uint64_t skeys[];   // [sort_cap] (last CTA only)
    __shared__ uint32_t red[2 * (kHistSelThreads / 32)];
    __shared__ uint32_t base_s[2], last_s;
    const uint32_t q = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t dk = p.thr[4 * q + 0], below = p.thr[4 * q + 1], keep_ties = p.thr[4 * q + 2];
    // exclusive prefix over the warp ranges of all preceding CTAs, then over the warps of this one
    const uint2* cnt = p.cnt + (size_t)q * p.sel_ctas * 8;
    uint32_t sl = 0, st = 0;
    for (uint32_t i = tid; i < blockIdx.x * 8u; i += kHistSelThreads) {
        const uint2 c = __ldcg(cnt + i);
        sl += c.x;
        st += c.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
        st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    if (lane == 0) {
        red[2 * warp] = sl;
        red[2 * warp + 1] = st;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t a = 0, b = 0;
        for (int w = 0; w < kHistSelThreads / 32; ++w) {
            a += red[2 * w];
            b += red[2 * w + 1];
        }
        base_s[0] = a;
        base_s[1] = b;
    }
    __syncthreads();
    uint32_t pos_l = base_s[0], pos_t = base_s[1];
    for (int w = 0; w < warp; ++w) {
        const uint2 c = __ldcg(cnt + (size_t)blockIdx.x * 8 + w);
        pos_l += c.x;
        pos_t += c.y;
    }
    const uint2 mine = __ldcg(cnt + (size_t)blockIdx.x * 8 + warp);
    uint64_t* stage = p.stage + (size_t)q * p.k;
    if (mine.x != 0 || (mine.y != 0 && pos_t < keep_ties)) {
        const uint16_t* d = p.dist + (size_t)q * p.n_pad;
        const unsigned long long r0 =
            (unsigned long long)blockIdx.x * kHistCtaRows + (unsigned long long)warp * kHistWarpRows;
        for (uint32_t i = 0; i < kHistWarpRows; i += 256) {
            const unsigned long long row = r0 + i + lane * 8u;
            uint32_t less, tie;
            hist_flags(d, row, p.N, dk, less, tie);
            if (!__any_sync(0xffffffffu, (less | tie) != 0u)) continue;
            // exclusive counts over the lanes before me (rows are lane-major inside the 256-row group)
            uint32_t packed = __popc(less) | (__popc(tie) << 16);
            uint32_t inc = packed;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            uint32_t il = pos_l + ((inc - packed) & 0xFFFFu), it = pos_t + ((inc - packed) >> 16);
            if (less | tie) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(d + row));
                const uint16_t* e = reinterpret_cast<const uint16_t*>(&v);
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const uint64_t key = ((uint64_t)e[x] << kGlobalRowBits) | (p.row_base + row + x);
                    if ((less >> x) & 1u) {
                        if (il < p.k) stage[il] = key;
                        ++il;
                    } else if ((tie >> x) & 1u) {
                        if (it < keep_ties) stage[below + it] = key;
                        ++it;
                    }
                }
            }
            pos_l += tot & 0xFFFFu;
            pos_t += tot >> 16;
        }
    }
    // ---- last CTA of the query: sort the k selected keys, write the outputs
    __threadfence();
    __syncthreads();
    if (tid == 0) last_s = atomicAdd(p.done + q, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (last_s == 0u) return;
    __threadfence();
    if (tid == 0) p.done[q] = 0;
    const uint32_t S = p.sort_cap;
    for (uint32_t i = tid; i < S; i += kHistSelThreads) skeys[i] = i < p.k ? __ldcg(stage + i) : kNone;
    __syncthreads();
    for (uint32_t size = 2; size <= S; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < (S >> 1); i += kHistSelThreads) {
                const uint32_t pos = 2 * i - (i & (stride - 1));
                const uint64_t a = skeys[pos], b = skeys[pos + stride];
                if ((a > b) == ((pos & size) == 0)) {
                    skeys[pos] = b;
                    skeys[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < p.k; i += kHistSelThreads) {
        const uint64_t key = skeys[i];
        const size_t o = (size_t)q * p.k + i;
        if (p.out_keys) p.out_keys[o] = key;
        if (p.out_rows) p.out_rows[o] = key == kNone ? kNone : (key & ((1ull << kGlobalRowBits) - 1ull));
        if (p.out_dists) p.out_dists[o] = key == kNone ? kNone : (key >> kGlobalRowBits);
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

HistPlan hist_plan(unsigned long long N, uint32_t W, uint32_t Q, uint32_t k, int sm_count, int smem_optin) {
    HistPlan pl = {};
    pl.applicable = false;
    if (W == 0 || W > kHistMaxWords || Q == 0 || k == 0 || k > 2048 || N == 0) return pl;
    if (N >= (1ull << 38)) return pl;
    if (sm_count <= 0) sm_count = 148;
    if (smem_optin <= 0) smem_optin = 232448;
    pl.bins = (64u * W + 1u + 31u) & ~31u;
    const bool spec = W <= 32;
    if (spec) {
        pl.tile_rows = kConsumerThreads * rows_per_thread((int)W);
    } else {
        uint32_t rows = (40u * 1024u) / (W * 8u);
        rows &= ~1u;
        if (rows > (uint32_t)kConsumerThreads) rows = kConsumerThreads;
        pl.tile_rows = rows;
    }
    const uint32_t tile_bytes = pl.tile_rows * W * 8;
    const uint32_t stride = (tile_bytes + 127u) & ~127u;
    const uint32_t per_q = hist_q_stride(W) + pl.bins * 4;
    uint32_t stages = 0, q_tile = 0;
    for (uint32_t S = 4; S >= 2; --S) {
        const int64_t avail = (int64_t)smem_optin - (int64_t)S * stride - (int64_t)S * 16 - 64;
        if (avail < (int64_t)per_q) continue;
        uint32_t qmax = (uint32_t)(avail / per_q);
        if (qmax > 32) qmax = 32;
        if (qmax >= (Q < 8 ? Q : 8) || S == 2) {
            stages = S;
            q_tile = qmax;
            break;
        }
    }
    if (!stages || !q_tile) return pl;
    pl.stages = stages;
    if (q_tile > Q) q_tile = Q;
    pl.num_qtiles = (Q + q_tile - 1) / q_tile;
    pl.q_tile = (Q + pl.num_qtiles - 1) / pl.num_qtiles;
    const unsigned long long tiles = (N + pl.tile_rows - 1) / pl.tile_rows;
    // chunks: ~4 waves of work items over the SMs, each at least 4 tiles
    unsigned long long want_chunks = ((unsigned long long)sm_count * 4 + pl.num_qtiles - 1) / pl.num_qtiles;
    unsigned long long tpc = (tiles + want_chunks - 1) / want_chunks;
    if (tpc < 4) tpc = 4;
    if (tpc > tiles) tpc = tiles;
    const unsigned long long chunks = (tiles + tpc - 1) / tpc;
    if (tpc * pl.tile_rows > 0xFFFFFFFFull || chunks * pl.num_qtiles > 0xFFFFFFFFull) return pl;
    pl.chunk_rows = (uint32_t)(tpc * pl.tile_rows);
    pl.num_chunks = (uint32_t)chunks;
    const unsigned long long items = chunks * pl.num_qtiles;
    pl.grid = (uint32_t)(items < (unsigned long long)sm_count ? items : (unsigned long long)sm_count);
    pl.smem = hist_smem_layout(tile_bytes, stages, W, pl.q_tile, pl.bins).total;
    if (pl.smem > (uint32_t)smem_optin) return pl;
    pl.n_pad = (N + 7ull) & ~7ull;
    const unsigned long long sel = (N + kHistCtaRows - 1) / kHistCtaRows;
    if (sel > 65535ull * 16) return pl;
    pl.sel_ctas = (uint32_t)sel;
    size_t off = 0;
    pl.off_hist = off;
    off += align256((size_t)Q * pl.bins * 4);
    pl.off_done = off;
    off += align256((size_t)Q * 4);
    pl.zero_bytes = off;                       // histograms + done counters must be zero on entry
    pl.off_thr = off;
    off += align256((size_t)Q * 16);
    pl.off_cnt = off;
    off += align256((size_t)Q * pl.sel_ctas * 8 * sizeof(uint2));
    pl.off_stage = off;
    off += align256((size_t)Q * k * 8);
    pl.off_dist = off;
    off += align256((size_t)Q * pl.n_pad * 2);
    pl.total = off;
    pl.applicable = true;
    return pl;
}

cudaError_t run_hist_top_n(const HistPlan& pl, const uint64_t* d_hashes, unsigned long long N, uint32_t W,
                           const uint64_t* d_queries, uint32_t Q, uint32_t k, unsigned long long row_base,
                           uint64_t* d_rows, uint64_t* d_dists, uint64_t* d_keys, void* ws, cudaStream_t stream,
                           cudaEvent_t* events) {
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    cudaError_t e = cudaMemsetAsync(w8, 0, pl.zero_bytes, stream);
    if (e != cudaSuccess) return e;
    HistParams hp;
    hp.hashes = d_hashes;
    hp.queries = d_queries;
    hp.dist = reinterpret_cast<uint16_t*>(w8 + pl.off_dist);
    hp.ghist = reinterpret_cast<uint32_t*>(w8 + pl.off_hist);
    hp.N = N;
    hp.n_pad = pl.n_pad;
    hp.Q = Q;
    hp.words = W;
    hp.bins = pl.bins;
    hp.chunk_rows = pl.chunk_rows;
    hp.num_chunks = pl.num_chunks;
    hp.q_tile = pl.q_tile;
    hp.num_qtiles = pl.num_qtiles;
    hp.stages = pl.stages;
    hp.tile_rows = pl.tile_rows;
    if (events && (e = cudaEventRecord(events[0], stream)) != cudaSuccess) return e;
    if (W <= 32) {
        HistLaunchFn fn = hist_launcher_for_width((int)W);
        if (!fn) return cudaErrorInvalidValue;
        e = fn(hp, (int)pl.grid, pl.smem, stream);
    } else {
        e = launch_hist_scan_generic(hp, (int)pl.grid, pl.smem, stream);
    }
    if (e != cudaSuccess) return e;
    if (events && (e = cudaEventRecord(events[1], stream)) != cudaSuccess) return e;
    HistSelParams sp;
    sp.dist = hp.dist;
    sp.ghist = hp.ghist;
    sp.thr = reinterpret_cast<uint32_t*>(w8 + pl.off_thr);
    sp.done = reinterpret_cast<uint32_t*>(w8 + pl.off_done);
    sp.cnt = reinterpret_cast<uint2*>(w8 + pl.off_cnt);
    sp.stage = reinterpret_cast<uint64_t*>(w8 + pl.off_stage);
    sp.out_rows = d_rows;
    sp.out_dists = d_dists;
    sp.out_keys = d_keys;
    sp.N = N;
    sp.n_pad = pl.n_pad;
    sp.row_base = row_base;
    sp.Q = Q;
    sp.k = k;
    sp.bins = pl.bins;
    sp.sel_ctas = pl.sel_ctas;
    sp.sort_cap = pow2_ceil(k < 2 ? 2 : k);
    const dim3 grid(pl.sel_ctas, Q);
    hist_count_kernel<<<grid, kHistSelThreads, 0, stream>>>(sp);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const size_t smem = (size_t)sp.sort_cap * 8;
    hist_emit_kernel<<<grid, kHistSelThreads, smem, stream>>>(sp);
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (events && (e = cudaEventRecord(events[2], stream)) != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace nps
