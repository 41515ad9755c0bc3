// ref_scan.cu — K2R host side + R0 (prefix kernel).  See ref_scan.cuh for the design.
#include <cmath>

#define NPS_REF_DEVICE_CODE 1
#include "ref_scan.cuh"

namespace nps {

// R0: distances of the first P rows of a query's table -> workspace (uint16), one row per thread, and
// tau_P = their k-th smallest (all-ones if P < k): every CTA adds its shared-memory histogram to the
// query's global one, the LAST CTA to finish turns it into tau_P (and zeroes it again).  With P == N
// (small tables) that CTA also replays the queue: the whole call is this one launch.
__global__ void __launch_bounds__(kRefPrefixThreads, 1) ref_prefix_kernel(const RefParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem);                     // [kRefBins]
    uint64_t* q_s = reinterpret_cast<uint64_t*>(smem + kRefBins * 4);       // [64]
    uint32_t* warp_sum = reinterpret_cast<uint32_t*>(q_s + 64);             // [32] + [1] "last" flag
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t W = p.words, P = p.P, qi = blockIdx.y;
    uint8_t* blk = p.ws + (size_t)qi * p.ws_stride;
    uint32_t* ctrl = p.ctrl + 4 * qi;
    uint32_t* ghist = p.hist + (size_t)qi * kRefBins;
    uint16_t* pre = reinterpret_cast<uint16_t*>(blk + p.off_pre);
    for (uint32_t i = tid; i < kRefBins; i += kRefPrefixThreads) hist[i] = 0;
    if ((uint32_t)tid < W) q_s[tid] = p.queries[(size_t)qi * W + tid];
    if (tid == 0 && blockIdx.x == 0) {     // R1 of this call starts only after every CTA of this grid
        ctrl[kRefCtrlOverflow] = 0;
        if (qi == 0) *p.any_overflow = 0;
    }
    __syncthreads();
    const uint32_t row = blockIdx.x * kRefPrefixThreads + tid;
    if (row < P) {
        const uint64_t* h = p.hashes + (size_t)row * W;
        uint32_t d = 0;
        if ((W & 1u) == 0u) {
            for (uint32_t w = 0; w < W; w += 2) {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(h + w));
                d += __popcll(v.x ^ q_s[w]) + __popcll(v.y ^ q_s[w + 1]);
            }
        } else {
            for (uint32_t w = 0; w < W; ++w) d += __popcll(__ldg(h + w) ^ q_s[w]);
        }
        pre[row] = (uint16_t)d;
        atomicAdd(&hist[d], 1u);
    }
    __syncthreads();
    for (uint32_t b = tid; b < kRefBins; b += kRefPrefixThreads) {
        const uint32_t c = hist[b];
        if (c) atomicAdd(&ghist[b], c);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t prev = atomicAdd(ctrl + kRefCtrlPrefixDone, 1u);
        warp_sum[32] = prev == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (warp_sum[32] == 0u) return;
    __threadfence();
    // ---- last CTA: k-th smallest distance; thread t owns bins [BPT * t, BPT * (t + 1))
    constexpr int BPT = kRefBins / kRefPrefixThreads;   // 8
    constexpr int NW = kRefPrefixThreads / 32;          // warps
    static_assert(BPT % 4 == 0 && NW <= 32, "histogram ownership");
    uint32_t b[BPT];
#pragma unroll
    for (int v = 0; v < BPT / 4; ++v) {
        const uint4 x = __ldcg(reinterpret_cast<const uint4*>(ghist) + tid * (BPT / 4) + v);
        b[4 * v + 0] = x.x;
        b[4 * v + 1] = x.y;
        b[4 * v + 2] = x.z;
        b[4 * v + 3] = x.w;
        *(reinterpret_cast<uint4*>(ghist) + tid * (BPT / 4) + v) = make_uint4(0u, 0u, 0u, 0u);   // clean for the next call
    }
    if (tid == 0) ctrl[kRefCtrlPrefixDone] = 0;
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < BPT; ++j) local += b[j];
    uint32_t inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < NW ? warp_sum[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        warp_sum[lane] = s;   // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = (warp ? warp_sum[warp - 1] : 0u) + inc - local;   // rows in bins below mine
    if (tid == 0 && warp_sum[31] < p.k) ctrl[kRefCtrlTau] = 0xFFFFFFFFu;      // fewer than k rows: no bound
    if (before < p.k && before + local >= p.k) {
        uint32_t c = before, tau = BPT * tid;
#pragma unroll
        for (int j = 0; j < BPT; ++j) {
            if (c < p.k && c + b[j] >= p.k) tau = BPT * tid + j;
            c += b[j];
        }
        ctrl[kRefCtrlTau] = tau;
    }
    if (P == p.N) {   // the whole table is prefix: replay here, no scan launch
        __syncthreads();
        replay_run(replay_args(p, qi), smem, tid, kRefPrefixThreads, 1);
    }
}

static uint32_t pow2_ceil_u32(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
static uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

RefPlan ref_plan(unsigned long long N, uint32_t W, uint32_t k, int sm_count, int smem_optin) {
    RefPlan pl = {};
    pl.applicable = false;
    if (W == 0 || W > kRefMaxWords || k == 0 || k > 2048) return pl;
    if (sm_count <= 0) sm_count = 148;
    if (smem_optin <= 0) smem_optin = 232448;
    if (N >= (1ull << 40)) return pl;
    const uint32_t prefix_base = kRefBins * 4 + 64 * 8 + 33 * 4 + 12;
    if (N <= kRefSmallRows) {
        pl.P = (uint32_t)N;
        pl.num_chunks = 0;
        pl.cap = 0;
        pl.chunk_tiles = 0;
        pl.tile_rows = 0;
        pl.stages = 0;
        pl.grid = 0;
    } else {
        double want = std::sqrt((double)N * (double)k);
        uint32_t P = want > (double)kRefMaxPrefix ? kRefMaxPrefix : (uint32_t)want;
        if (P < 2 * k) P = 2 * k;
        if (P < 1024) P = 1024;
        P = round_up(P, 256);
        if (P > kRefMaxPrefix) P = kRefMaxPrefix;
        pl.P = P;
        const bool spec = W <= 32;
        if (spec) {
            pl.tile_rows = kConsumerThreads * rows_per_thread((int)W);
        } else {
            uint32_t rows = (40u * 1024u) / (W * 8u);
            rows &= ~1u;
            if (rows > (uint32_t)kConsumerThreads) rows = kConsumerThreads;
            pl.tile_rows = rows;
        }
        const unsigned long long rows1 = N - P;
        const unsigned long long tiles = (rows1 + pl.tile_rows - 1) / pl.tile_rows;
        const double per_tile = (double)pl.tile_rows * k / P;       // expected survivors per tile
        unsigned long long ct = (tiles + sm_count - 1) / sm_count;   // one chunk per CTA if the segment fits
        const double max_ct = ((double)kRefMaxCap - 64.0) / (4.0 * per_tile);
        if ((double)ct > max_ct) ct = max_ct < 1.0 ? 1 : (unsigned long long)max_ct;
        if (ct < 1) ct = 1;
        const unsigned long long chunks = (tiles + ct - 1) / ct;
        if (chunks > kRefMaxChunks || ct > 0xFFFFFFFFull) return pl;
        pl.chunk_tiles = (uint32_t)ct;
        pl.num_chunks = (uint32_t)chunks;
        uint32_t cap = pow2_ceil_u32((uint32_t)(4.0 * per_tile * (double)ct + 64.0));
        if (cap < 64) cap = 64;
        if (cap > kRefMaxCap) cap = kRefMaxCap;
        pl.cap = cap;
        pl.grid = (uint32_t)(chunks < (unsigned long long)sm_count ? chunks : (unsigned long long)sm_count);
        const uint32_t tile_bytes = pl.tile_rows * W * 8;
        uint32_t stages = 0;
        for (uint32_t S = 4; S >= 2; --S)
            if (ref_smem_layout(tile_bytes, S, W, cap, k, pl.num_chunks).total <= (uint32_t)smem_optin) {
                stages = S;
                break;
            }
        if (!stages) return pl;
        pl.stages = stages;
        pl.smem_scan = ref_smem_layout(tile_bytes, stages, W, cap, k, pl.num_chunks).total;
    }
    const uint32_t replay = ref_replay_smem(k, pl.num_chunks);
    pl.smem_prefix = prefix_base > replay ? prefix_base : replay;
    if (pl.smem_prefix > (uint32_t)smem_optin) return pl;
    size_t off = 0;
    pl.off_pre = (uint32_t)off;
    off += ((size_t)pl.P * 2 + 15) & ~(size_t)15;
    pl.off_cnt = (uint32_t)off;
    off += ((size_t)pl.num_chunks * 4 + 15) & ~(size_t)15;
    pl.off_seg = (uint32_t)off;
    off += (size_t)pl.num_chunks * pl.cap * 8;
    if (off > 0xFFFFFFFFull) return pl;
    pl.ws_stride = (off + 255) & ~(size_t)255;
    pl.applicable = true;
    return pl;
}

cudaError_t run_ref_top_n(const RefPlan& pl, const uint64_t* d_hashes, unsigned long long N, uint32_t W,
                          const uint64_t* d_queries, uint32_t nq, uint32_t k, uint64_t* d_rows, uint64_t* d_scores,
                          void* ws, bool zero_first, cudaStream_t stream) {
    const size_t hdr = ((size_t)16 + (size_t)nq * 16 + 255) & ~(size_t)255;
    cudaError_t e;
    if (zero_first) {
        e = cudaMemsetAsync(ws, 0, ref_zero_bytes(nq), stream);
        if (e != cudaSuccess) return e;
    }
    RefParams p;
    p.hashes = d_hashes;
    p.queries = d_queries;
    p.out_rows = d_rows;
    p.out_scores = d_scores;
    p.any_overflow = reinterpret_cast<uint32_t*>(ws);
    p.ctrl = reinterpret_cast<uint32_t*>(ws) + 4;
    p.hist = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ws) + hdr);
    p.ws = reinterpret_cast<uint8_t*>(ws) + ref_zero_bytes(nq);
    p.ws_stride = pl.ws_stride;
    p.N = N;
    p.words = W;
    p.k = k;
    p.P = pl.P;
    p.tile_rows = pl.tile_rows;
    p.chunk_tiles = pl.chunk_tiles;
    p.num_chunks = pl.num_chunks;
    p.cap = pl.cap;
    p.stages = pl.stages;
    p.off_pre = pl.off_pre;
    p.off_cnt = pl.off_cnt;
    p.off_seg = pl.off_seg;
    e = cudaFuncSetAttribute(ref_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_prefix);
    if (e != cudaSuccess) return e;
    const dim3 pgrid(pl.P ? (pl.P + kRefPrefixThreads - 1) / kRefPrefixThreads : 1u, nq);
    ref_prefix_kernel<<<pgrid, kRefPrefixThreads, pl.smem_prefix, stream>>>(p);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess || pl.num_chunks == 0) return e;
    const dim3 grid(pl.grid, nq);
    if (W <= 32) {
        RefLaunchFn fn = ref_launcher_for_width((int)W);
        if (!fn) return cudaErrorInvalidValue;
        return fn(p, grid, pl.smem_scan, stream);
    }
    return launch_ref_scan_generic(p, grid, pl.smem_scan, stream);
}

}  // namespace nps
