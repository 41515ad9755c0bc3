// mma_scan.cu — K2T kernels: tensor-core batched Hamming scan + candidate select.
// See mma_scan.cuh for the design.  Replaces, for query batches, the scan bodies of
// np_sims/ufunc.c:207-441 (XOR + popcount + queue insert) with an exact +-1 GEMM on E2M1 operands
// (tcgen05 kind::mxf4, unit block scales).
#include <cstdio>
#include <cstdlib>

#include "mma_scan.cuh"

namespace nps {

// Bit-packed signature word -> E2M1 operand words, without spreading any bits.
// An E2M1 nibble is  sign | exp1 exp0 | mant ; 0b0010 = +1.0, 0b1010 = -1.0.  Every nibble of the
// raw 32-bit word already carries four signature bits, so  ((R << s) & 0x88888888) | 0x22222222
// turns bit (3 - s) of every nibble into the sign of a +-1.0 element: four shifts give four operand
// words holding all 32 bits.  The K order of the elements is a fixed permutation of the bit order,
// the same for signatures and queries, so dot products (= bits - 2 * Hamming distance) are unchanged.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Expand 16-byte chunks [C0, C1) of one 256-bit K-block (words w[0..3], the first `nvalid` of which
// exist) of operand row `r` into its 128-byte swizzled row.  Chunk c holds signature bits
// [32c, 32c+32) of the K-block.  tile_base: shared address of the [rows x 128 B] SWIZZLE_128B
// tile (1024-byte aligned).  Missing words become E2M1 zeros, which add nothing to the dot.
template <int C0, int C1>
__device__ __forceinline__ void expand_row_kblock(uint32_t tile_base, uint32_t r, const uint64_t (&w)[4],
                                                  uint32_t nvalid) {
    const uint32_t r7 = r & 7u;
    const uint32_t row_base = tile_base + (r >> 3) * 1024u + r7 * 128u;
#pragma unroll
    for (int c = C0; c < C1; ++c) {
        const uint32_t h = (c & 1) ? (uint32_t)(w[c >> 1] >> 32) : (uint32_t)w[c >> 1];
        uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;     // missing words: E2M1 zeros add nothing to the dot
        if ((uint32_t)(c >> 1) < nvalid) {
            x0 = (h & 0x88888888u) | 0x22222222u;
            x1 = ((h << 1) & 0x88888888u) | 0x22222222u;
            x2 = ((h << 2) & 0x88888888u) | 0x22222222u;
            x3 = ((h << 3) & 0x88888888u) | 0x22222222u;
        }
        sts128(row_base + (((uint32_t)c ^ r7) << 4), x0, x1, x2, x3);
    }
}

// Hot-path form for the A (row) tiles: exactly NW words of the K-block exist for every row of the
// tile (the issuer never reads the chunks of missing words, so they are not written at all), no
// per-chunk predicates, and (x & 0x8888...) | 0x2222... as ONE three-input LOP3 with the constants
// in registers (with two immediates ptxas emits two LOP3s, plus CS2R/BSSY/BSYNC for every chunk of
// the predicated form: 18 instructions per chunk measured, 9 here).
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t m, uint32_t o) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(m), "r"(o));
    return d;
}
template <int NW>
__device__ __forceinline__ void expand_row_words(uint32_t tile_base, uint32_t r, const uint64_t (&w)[4]) {
    const uint32_t r7 = r & 7u;
    const uint32_t a0 = (tile_base + (r >> 3) * 1024u + r7 * 128u) | (r7 << 4);   // chunk c lives at a0 ^ (c << 4)
    uint32_t m8, o2;
    asm volatile("mov.b32 %0, 0x88888888;" : "=r"(m8));
    asm volatile("mov.b32 %0, 0x22222222;" : "=r"(o2));
#pragma unroll
    for (int c = 0; c < 2 * NW; ++c) {
        const uint32_t h = (c & 1) ? (uint32_t)(w[c >> 1] >> 32) : (uint32_t)w[c >> 1];
        sts128(a0 ^ ((uint32_t)c << 4), and_or(h, m8, o2), and_or(h << 1, m8, o2), and_or(h << 2, m8, o2),
               and_or(h << 3, m8, o2));
    }
}

// ---- thresholds folded into the GEMM -------------------------------------------------------
// The filter "dot >= T" (T = bits - 2 * k-th best distance, an even integer) is evaluated by the
// tensor core: one extra K = 64 slice whose A elements are all 4.0 and whose B row n spells -T_n / 4
// as a sum of E2M1 values, so the accumulator holds  dot - T  and the epilogue only looks at sign
// bits.  4 * {0, .5, 1, 1.5, 2, 3, 4, 6} = {0, 2, 4, 6, 8, 12, 16, 24}: |T| / 2 = 12 * n12 + rem with
// rem < 12 written as at most two of {1, 2, 3, 4, 6, 8}.  One slice spells |T| <= 1510 (23 words);
// wider signatures use two slices, T = T1 + T2 with |T1|, |T2| <= 1510 (thr_split).
__device__ __forceinline__ float clamp_thr(float t, uint32_t words) {
    const float r = words > kMmaThrOneSlice ? (float)kMmaThrRange : (float)kMmaThrSlice;
    return fminf(fmaxf(t, -r), r);   // +-inf -> never / always
}
__device__ __forceinline__ void thr_split(float t, uint32_t words, int& T1, int& T2) {
    const int T = (int)clamp_thr(t, words);
    T1 = min(max(T, -kMmaThrSlice), kMmaThrSlice);
    T2 = T - T1;
}
__device__ __forceinline__ void encode_thr(int T, uint32_t (&out)[8]) {
    const uint32_t m = (uint32_t)(T < 0 ? -T : T) >> 1;
    const int n12 = (int)(m / 12u);
    const uint32_t rem = m - 12u * (uint32_t)n12;
    // E2M1 codes of the one or two remainder elements, packed as c1 | c2 << 4, indexed by rem
    uint32_t c1, c2;
    switch (rem) {
        case 0: c1 = 0; c2 = 0; break;
        case 1: c1 = 1; c2 = 0; break;
        case 2: c1 = 2; c2 = 0; break;
        case 3: c1 = 3; c2 = 0; break;
        case 4: c1 = 4; c2 = 0; break;
        case 5: c1 = 4; c2 = 1; break;
        case 6: c1 = 5; c2 = 0; break;
        case 7: c1 = 5; c2 = 1; break;
        case 8: c1 = 6; c2 = 0; break;
        case 9: c1 = 6; c2 = 1; break;
        case 10: c1 = 6; c2 = 2; break;
        default: c1 = 6; c2 = 3; break;
    }
    const uint32_t sign = T > 0 ? 0x88888888u : 0u;   // B holds -T: negative elements for positive T
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int full = min(max(n12 - 8 * w, 0), 8);
        uint32_t x = full == 8 ? 0x77777777u : (0x77777777u & ((1u << (4 * full)) - 1u));
        const int p1 = n12 - 8 * w, p2 = p1 + 1;
        if (p1 >= 0 && p1 < 8) x |= c1 << (4 * p1);
        if (p2 >= 0 && p2 < 8) x |= c2 << (4 * p2);
        out[w] = x | sign;
    }
}

__device__ __forceinline__ uint32_t phase_tile(const MmaScanParams& p, uint32_t i) {
    const uint32_t j = p.skip_mod ? (i + i / (p.skip_mod - 1u) + 1u) : i;
    return j * p.tile_stride;
}

__device__ __forceinline__ void global_append(const MmaScanParams& p, uint32_t q, uint64_t key) {
    const uint32_t pos = atomicAdd(p.cnt + q, 1u);
    if (pos < p.cap) p.cand[(size_t)q * p.cap + pos] = key;
}

// Slow path of the epilogue: some value  dot - T  of the group is non-negative, i.e. a row passed
// the distance filter (distance <= k-th best distance; ties are sorted out by mma_select_kernel).
// While the accumulator is still held, the survivors are only compacted into this warp's
// shared-memory staging buffer as (raw accumulator bits, lane | column << 8): the whole warp walks
// the columns of the flagged 8-column quarters together, one vote per column, slots handed out
// from a warp-uniform register counter — no atomics, no divergence, no calls (this code is cold in
// the instruction cache; a version with one out-of-line call per column cost ~4000 cycles per
// group in fetch stalls, one with a shared-memory atomic per survivor ~2000).  Decoding
// (distance, key) and the global append happen at flush time, after the accumulator went back to
// the MMA issuer.
template <int J>
__device__ __forceinline__ void stage_column(uint32_t x, uint32_t code, uint32_t& wn, uint2* hits, uint32_t lane_lt) {
    const uint32_t mask = __ballot_sync(0xffffffffu, (int)x >= 0);
    if (mask) {   // warp-uniform
        if ((int)x >= 0) {
            const uint32_t slot = min(wn + __popc(mask & lane_lt), (uint32_t)kMmaHitsPerWarp);
            hits[slot] = make_uint2(x, code + (uint32_t)(J << 8));
        }
        wn += __popc(mask);
    }
}
template <int J0, int I = 0>
__device__ __forceinline__ void stage_hits8(const uint32_t (&v)[56], bool mine, uint32_t code, uint32_t& wn, uint2* hits,
                                            uint32_t lane_lt) {
    if constexpr (I < 8) {
        stage_column<J0 + I>(mine ? v[J0 + I] : 0x80000000u, code, wn, hits, lane_lt);
        stage_hits8<J0, I + 1>(v, mine, code, wn, hits, lane_lt);
    }
}

// Exact fallback when a warp's staging buffer overflowed during a tile (early phases, whose
// thresholds are still loose): the warp's accumulator columns are read again and every survivor is
// appended directly.  Cold and slow by design (rolled loop, values parked in local memory).
__device__ __noinline__ void rescan_group(uint32_t taddr, const float* thr, uint32_t q_first, int ncols, uint64_t grow,
                                          int bits, bool valid, uint32_t* g_cnt, uint64_t* g_cand, uint32_t cap) {
    uint32_t v[32];
    tmem_ld32(taddr, v);      // .sync.aligned: the whole warp is here, converged
    tmem_ld_wait();
    if (valid) {
#pragma unroll 1
        for (int j = 0; j < ncols; ++j) {
            const uint32_t x = v[j];
            if ((int)x >= 0) {
                const int dot = (int)(__uint_as_float(x) + thr[j]);
                const uint64_t key = ((uint64_t)((bits - dot) >> 1) << kGlobalRowBits) | grow;
                const uint32_t pos = atomicAdd(g_cnt + q_first + j, 1u);
                if (pos < cap) g_cand[(size_t)(q_first + j) * cap + pos] = key;
            }
        }
    }
    __syncwarp();
}

// development aid: CTA 0 logs (clock, role, event) triples; role slots of 4096 events each
#ifdef NPS_MMA_DEV
#define NPS_TRACE(role, ev)                                                                        \
    do {                                                                                           \
        if (p.trace && blockIdx.x == 0 && trace_n < 4096u) {                                       \
            p.trace[((role) * 4096u + trace_n) * 2u] = (unsigned long long)clock64();              \
            p.trace[((role) * 4096u + trace_n) * 2u + 1u] = (ev);                                  \
            ++trace_n;                                                                             \
        }                                                                                          \
    } while (0)
#define NPS_DEV(...) __VA_ARGS__
#else
#define NPS_TRACE(role, ev) do { } while (0)
#define NPS_DEV(...)
#endif

// Warp roles, in order of warp id: epilogue [0, 16), expanders [16, 24), producer, two MMA issuers,
// one idle warp.  The hardware arbiter favours the highest warp id of a scheduler, so the roles
// whose latency paces the pipeline come last: the single-thread issuers, then the expanders
// (measured: with the expanders on the lowest ids, a tile's expansion took ~1900 cycles for ~150
// instructions per thread because the waiting epilogue warps kept winning the issue slots).
constexpr int kWarpProducer = kMmaExpWarps + kMmaEpiWarps;
constexpr int kWarpMma = kWarpProducer + 1;      // two issuer warps: kWarpMma (even tiles), kWarpMma + 1 (odd tiles)

// Register redistribution between warpgroups (setmaxnreg): the kernel launches with 96 registers
// per thread; expanders and the single-thread roles give registers back, the epilogue takes them.
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int kRegsExpander = 64, kRegsMisc = 32;
constexpr int kRegsLaunch = (65536 / kMmaThreads) / 8 * 8;   // what __launch_bounds__(kMmaThreads, 1) allows
constexpr int kRegsEpilogue =
    ((kMmaThreads * kRegsLaunch - kMmaExpWarps * 32 * kRegsExpander - 4 * 32 * kRegsMisc) / (kMmaEpiWarps * 32)) / 8 * 8;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// kPair: the CTA is one half of a CTA pair (cluster of 2 on one TPC).  The pair issues ONE
// tcgen05.mma.cta_group::2 of M = 256 per signature word: this CTA supplies the 128 rows of ITS row tile
// (tile 2j + rank of the phase) and HALF of the query tile (B rows [rank * NT/2, (rank + 1) * NT/2)) from
// its own shared memory, and keeps its 128 accumulator rows x NT columns in its own TMEM.  Per MMA a CTA's
// shared memory then serves 4 KB of A + 3.5 KB of B instead of 4 + 7 KB — the single-CTA kernel is bound by
// exactly that traffic.  Only the leader (rank 0) issues; barriers that gate the issue
// (A group full, accumulator free, query tile ready) live in the LEADER's shared memory and take
// arrivals from both CTAs; barriers signalled by tcgen05.commit are multicast to both.
template <bool kDense, bool kPair>
__global__ void __launch_bounds__(kMmaThreads, 1) mma_scan_kernel(const MmaScanParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t W = p.words, KB = mma_kblocks(W), NT = p.n_tile;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const uint32_t cta_id = kPair ? (blockIdx.x >> 1) : blockIdx.x;       // work is dealt to pairs
    const uint32_t num_ctas = kPair ? (gridDim.x >> 1) : gridDim.x;
    const uint32_t NB = kPair ? NT / 2u : NT;                             // B rows (queries) this CTA expands
    const uint32_t qoff = rank * NB;                                      // ... starting at this column of the tile
    const uint32_t ptiles = kPair ? (p.phase_tiles + 1u) / 2u : p.phase_tiles;   // loop units (tile pairs / tiles)
    // QT = 2 (pairs only): a work item covers TWO query tiles.  Each CTA then keeps its halves of both in shared
    // memory (together what one CTA of the single-CTA kernel holds for one tile) and every expanded row tile
    // is multiplied with both — half as many expansions per MMA.  MMA issuer h owns query tile h of the item
    // and accumulator buffer h; an A slot is released when BOTH issuers' MMAs on it have retired.
    const uint32_t QT = kPair ? p.qt_per_item : 1u;
    const uint32_t num_qgroups = (p.num_qtiles + QT - 1u) / QT;
    const MmaSmem L = mma_smem_layout(W, NT, p.a_stages, p.p_stages, kPair ? 1u : 0u, QT);
    const uint32_t bthr_stride = NB * 32u * (W > kMmaThrOneSlice ? 2u : 1u);   // threshold operand tile of one query tile
    const uint32_t PS = p.p_stages;
    // A ring: each MMA issuer owns RG (1 or 2) groups of GS K-block stages, so that every
    // group barrier has exactly one consumer (mbarrier phase parity cannot alias).
    const uint32_t GS = p.group_kb, RG = p.ring_per_issuer, NG = 2u * RG;
    const uint32_t GPT = (KB + GS - 1) / GS;             // groups per row tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* p_full = bars;
    uint64_t* p_empty = p_full + PS;
    uint64_t* g_full = p_empty + PS;
    uint64_t* g_empty = g_full + NG;
    // acc_full: TWO barriers per accumulator buffer, used alternately by the tiles of that buffer
    // (tile m of buffer b -> acc_full[2 b + (m & 1)], phase m >> 1).  With two A groups per issuer the
    // expanders run up to two tiles ahead of the MMAs of a buffer; a single barrier per buffer would then
    // be waited on one phase ahead at item boundaries, which a parity wait reports as "done".
    uint64_t* acc_full = g_empty + NG;      // [4]
    uint64_t* acc_empty = acc_full + 4;     // [2]
    uint64_t* b_full = acc_empty + 2;       // [1]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + L.tmem_off);
    uint32_t* wcnt_s = reinterpret_cast<uint32_t*>(smem + L.wcnt_off);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < PS; ++s) {
            mbar_init(&p_full[s], 1);
            mbar_init(&p_empty[s], kMmaExpWarps / 2);    // one expander set per stage
        }
        constexpr uint32_t kCtas = kPair ? 2u : 1u;     // leader barriers count the arrivals of both CTAs
        for (uint32_t s = 0; s < NG; ++s) {
            mbar_init(&g_full[s], kCtas * (kMmaExpWarps / 2));     // one expander set per issuer (and CTA)
            mbar_init(&g_empty[s], QT);                            // one commit per issuer that read the group
        }
        for (int b = 0; b < 4; ++b) mbar_init(&acc_full[b], 1);
        for (int b = 0; b < 2; ++b) mbar_init(&acc_empty[b], kCtas * kMmaEpiWarps);
        mbar_init(b_full, kCtas * kMmaExpWarps);
        fence_mbar_init();
    }
    if (tid < kMmaEpiWarps) wcnt_s[tid] = 0;
    for (uint32_t i = tid; i < kMmaTileRows * 32u / 4u; i += kMmaThreads)
        reinterpret_cast<uint32_t*>(smem + L.athr_off)[i] = 0x66666666u;   // E2M1 4.0 in every element
    fence_proxy_async_smem();
    if (warp == kWarpMma) {
        if constexpr (kPair) tmem_alloc_pair(tmem_ptr_s, 512);
        else tmem_alloc(tmem_ptr_s, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_s);
    // Block scale factors: every byte of TMEM columns [kMmaSfCol, 512) = 0x7F = UE8M0 1.0, so the
    // (layout-dependent) scale lookup of kind::mxf4 returns 1.0 for every row and K block.
    if (warp < 4) {
        for (uint32_t c = kMmaSfCol; c < 512u; c += 8u)
            tmem_st8_fill(tmem_base + (((uint32_t)(warp & 3) * 32u) << 16) + c, 0x7F7F7F7Fu);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if constexpr (kPair) cluster_sync_all();   // both CTAs' barriers and tensor memory are set up

    // Everything above touched only this CTA's shared and tensor memory; from here on the kernel reads
    // what the previous launch in the stream (the select of the phase before) wrote.
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t total_items = p.num_chunks * num_qgroups;
    const uint32_t row_bytes = W * 8;
    NPS_DEV(uint32_t trace_n = 0;)

    if (warp == kWarpProducer) {
        // =========================== producer: bit-packed row tiles ===========================
        reg_dec<kRegsMisc>();
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t item = cta_id; item < total_items; item += num_ctas) {
                const uint32_t chunk = item / num_qgroups;
                const uint32_t i0 = chunk * p.tiles_per_chunk;
                const uint32_t i1 = min(i0 + p.tiles_per_chunk, ptiles);
                for (uint32_t i = i0; i < i1; ++i) {
                    const uint32_t ti = kPair ? 2u * i + rank : i;      // this CTA's row tile of the phase
                    const bool tile_ok = ti < p.phase_tiles;            // odd tile count: the last pair is half empty
                    const unsigned long long row0 = (unsigned long long)phase_tile(p, tile_ok ? ti : 0u) * kMmaTileRows;
                    const unsigned long long left = p.N - row0;
                    const uint32_t rows = !tile_ok ? 0u : left < (unsigned long long)kMmaTileRows ? (uint32_t)left : kMmaTileRows;
                    mbar_wait(&p_empty[s], ph ^ 1u);
                    const uint32_t bytes = rows * row_bytes, bulk = bytes & ~15u;
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.hashes) + row0 * row_bytes;
                    uint8_t* dst = smem + L.p_off + s * L.p_stride;
                    if (bytes != bulk)
                        *reinterpret_cast<volatile uint64_t*>(dst + bulk) =
                            *reinterpret_cast<const uint64_t*>(src + bulk);
                    mbar_arrive_expect_tx(&p_full[s], bulk);
                    if (bulk) bulk_g2s(dst, src, bulk, &p_full[s]);
                    if (++s == PS) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == kWarpMma + 2) {
        // idle warp: completes the last warpgroup so that setmaxnreg is warpgroup-uniform
        reg_dec<kRegsMisc>();
    } else if (warp >= kWarpMma) {
        // =========================== MMA issuers ================================================
        // Two warps, one per accumulator buffer (even / odd row tiles), each running converged
        // with one elected lane issuing: while one sits in a barrier wait or in the stall that
        // follows tcgen05.commit, the other keeps the tensor pipe fed.  Their MMA streams are
        // independent (different accumulators, different A groups), so no order is required.
        reg_dec<kRegsMisc>();
        const uint32_t mw = warp - kWarpMma;
        const uint32_t idesc = umma_idesc_mxf4(kPair ? 2u * kMmaTileRows : kMmaTileRows, NT);
        auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t sa, uint32_t sb, uint32_t acc) {
            if constexpr (kPair) umma_mxf4_pair(d, a, b, idesc, sa, sb, acc);
            else umma_mxf4(d, a, b, idesc, sa, sb, acc);
        };
        auto commit = [&](uint64_t* bar) {
            if constexpr (kPair) umma_commit_pair(bar);
            else umma_commit(bar);
        };
        auto wait_in = [&](uint64_t* bar, uint32_t parity) {   // barriers with arrivals from both CTAs
            if constexpr (kPair) mbar_wait_cluster(bar, parity);
            else mbar_wait(bar, parity);
        };
        const uint32_t sfa = tmem_base + kMmaSfCol, sfb = tmem_base + kMmaSfCol + 32u;
        const uint64_t a_desc0 = umma_desc_k_sw128(base + L.a_off);
        const uint64_t b_desc0 = umma_desc_k_sw128(base + L.b_off);
        const uint64_t athr_desc = umma_desc_k_noswizzle(base + L.athr_off, 128u, 256u);
        const uint64_t bthr_desc = umma_desc_k_noswizzle(base + L.bthr_off, 128u, 256u);
        const uint64_t bthr2_desc = umma_desc_k_noswizzle(base + L.bthr_off + NB * 32u, 128u, 256u);
        uint32_t t_idx = 0, it_idx = 0, lg = 0;   // lg: groups consumed by this issuer
        if (QT == 2u) {
            // Two query tiles per item: this issuer multiplies EVERY row tile of the item with query tile mw into
            // accumulator buffer mw (use index = tile count).  Plans with QT = 2 have one whole-tile A group per
            // expander set: slot = tile & 1.
            const uint64_t b_desc_h = b_desc0 + (uint64_t)((mw * KB * NB * kMmaKBlockBytes) >> 4);
            const uint64_t bthr_h = umma_desc_k_noswizzle(base + L.bthr_off + mw * bthr_stride, 128u, 256u);
            const uint64_t bthr2_h = umma_desc_k_noswizzle(base + L.bthr_off + mw * bthr_stride + NB * 32u, 128u, 256u);
            const uint32_t d_tmem = tmem_base + mw * NT;
            for (uint32_t item = cta_id; item < total_items && rank == 0u; item += num_ctas, ++it_idx) {
                const uint32_t chunk = item / num_qgroups;
                const uint32_t i0 = chunk * p.tiles_per_chunk;
                const uint32_t i1 = min(i0 + p.tiles_per_chunk, ptiles);
                if (i0 < i1) wait_in(b_full, it_idx & 1u);
                for (uint32_t i = i0; i < i1; ++i, ++t_idx) {
                    wait_in(&acc_empty[mw], (t_idx & 1u) ^ 1u);
                    tc_fence_after();
                    if (lane == 0) NPS_TRACE(mw ? 3 : 0, 1);   // accumulator free
                    const uint32_t slot = t_idx & 1u;
                    wait_in(&g_full[slot], (t_idx >> 1) & 1u);
                    tc_fence_after();
                    if (lane == 0) NPS_TRACE(mw ? 3 : 0, 2);   // A group full
                    if (elect_one()) {
                        for (uint32_t kb = 0; kb < KB; ++kb) {
                            const uint64_t a_desc = a_desc0 + (uint64_t)(((slot * GS + kb) * kMmaAStageBytes) >> 4);
                            const uint64_t b_desc = b_desc_h + (uint64_t)((kb * NB * kMmaKBlockBytes) >> 4);
                            const uint32_t nw = min(4u, W - 4u * kb);
                            for (uint32_t j = 0; j < nw && !(p.debug & 4u); ++j)
                                mma(d_tmem, a_desc + 2u * j, b_desc + 2u * j, sfa, sfb, (kb | j) != 0u);
                        }
                        commit(&g_empty[slot]);          // the slot is free once BOTH issuers' MMAs on it have retired
                        if (!p.dense) {
                            mma(d_tmem, athr_desc, bthr_h, sfa, sfb, 1u);
                            if (W > kMmaThrOneSlice) mma(d_tmem, athr_desc, bthr2_h, sfa, sfb, 1u);
                        }
                        commit(&acc_full[2u * mw + (t_idx & 1u)]);
                    }
                    __syncwarp();
                    if (lane == 0) NPS_TRACE(mw ? 3 : 0, 3);   // issued + committed
                }
            }
        } else
        for (uint32_t item = cta_id; item < total_items && rank == 0u; item += num_ctas, ++it_idx) {
            const uint32_t chunk = item / num_qgroups;
            const uint32_t i0 = chunk * p.tiles_per_chunk;
            const uint32_t i1 = min(i0 + p.tiles_per_chunk, ptiles);
            bool b_ready = false;
            for (uint32_t i = i0; i < i1; ++i, ++t_idx) {
                const uint32_t buf = t_idx & 1u;
                if (buf != mw) continue;
                if (!b_ready) {
                    wait_in(b_full, it_idx & 1u);
                    b_ready = true;
                }
                wait_in(&acc_empty[buf], ((t_idx >> 1) & 1u) ^ 1u);
                tc_fence_after();
                if (lane == 0) NPS_TRACE(mw ? 3 : 0, 1);   // accumulator free
                const uint32_t d_tmem = tmem_base + buf * NT;
                for (uint32_t g = 0; g < GPT; ++g, ++lg) {
                    const uint32_t slot = mw * RG + lg % RG, par = (lg / RG) & 1u;
                    wait_in(&g_full[slot], par);
                    tc_fence_after();
                    if (lane == 0) NPS_TRACE(mw ? 3 : 0, 2);   // A group full
                    if (elect_one()) {
                        const uint32_t kb0 = g * GS, kb1 = min(kb0 + GS, KB);
                        for (uint32_t kb = kb0; kb < kb1; ++kb) {
                            const uint32_t stage = slot * GS + (kb - kb0);
                            const uint64_t a_desc = a_desc0 + (uint64_t)((stage * kMmaAStageBytes) >> 4);
                            const uint64_t b_desc = b_desc0 + (uint64_t)((kb * NB * kMmaKBlockBytes) >> 4);
                            const uint32_t nw = min(4u, W - 4u * kb);   // one MMA (K = 64) per signature word
                            for (uint32_t j = 0; j < nw && !(p.debug & 4u); ++j)
                                mma(d_tmem, a_desc + 2u * j, b_desc + 2u * j, sfa, sfb, (kb | j) != 0u);
                        }
                        commit(&g_empty[slot]);                       // group reusable once these MMAs retire
                        if (g + 1 == GPT) {
                            if (!p.dense) {                                                                // ... - T
                                mma(d_tmem, athr_desc, bthr_desc, sfa, sfb, 1u);
                                if (W > kMmaThrOneSlice) mma(d_tmem, athr_desc, bthr2_desc, sfa, sfb, 1u);
                            }
                            commit(&acc_full[2u * buf + ((t_idx >> 1) & 1u)]);   // accumulator complete -> epilogue
                        }
                    }
                    __syncwarp();
                    if (lane == 0) NPS_TRACE(mw ? 3 : 0, 3);   // issued + committed
                }
            }
        }
    } else if (warp >= kMmaEpiWarps) {
        // =========================== expanders: bits -> E2M1 operand tiles =======================
        // kMmaExpWarps/4 threads per operand row; with two, each expands half of the 128-byte row
        reg_dec<kRegsExpander>();
        constexpr int kParts = kMmaExpWarps / 4;
        const uint32_t u = (tid - kMmaEpiWarps * 32) & 127u;   // operand row
        const uint32_t part = (tid - kMmaEpiWarps * 32) >> 7;  // which slice of the 128-byte row
        auto expand = [&](uint32_t tile_base, uint32_t r, const uint64_t (&w)[4], uint32_t nvalid) {
            if constexpr (kParts == 1) {
                expand_row_kblock<0, 8>(tile_base, r, w, nvalid);
            } else {
                if (part == 0) expand_row_kblock<0, 4>(tile_base, r, w, nvalid);
                else expand_row_kblock<4, 8>(tile_base, r, w, nvalid);
            }
        };
        static_assert(kMmaExpWarps == 8, "two expander sets of four warps");
        auto arrive_in = [&](uint64_t* bar) {   // barriers the MMA issuer (leader CTA) waits on
            if constexpr (kPair) mbar_arrive_leader(bar);
            else mbar_arrive(bar);
        };
        // packed ring: PS (even) stages; my set's j-th tile sits in stage part + 2 * (j % (PS / 2)) — the producer
        // fills stage t % PS for tile t, and t % 2 is the set
        const uint32_t PH = PS >> 1;
        uint32_t pcur = 0, pph = 0, t_idx = 0, lg2[1] = {0};   // pcur: my set's position in its half of the ring, pph: its parity
        for (uint32_t item = cta_id; item < total_items; item += num_ctas) {
            const uint32_t chunk = item / num_qgroups, qg = item - chunk * num_qgroups;
            const uint32_t i0 = chunk * p.tiles_per_chunk;
            const uint32_t i1 = min(i0 + p.tiles_per_chunk, ptiles);
            // B (query tile) is resident for the whole item; the previous item's MMAs must have
            // retired before it is overwritten.
            // Each expander set waits for ITS OWN buffer's last tile only — it has just expanded that tile,
            // so the barrier cannot be more than one phase behind and the parity test is unambiguous (a set
            // waiting on the other, possibly lagging, set's barrier can be two phases ahead of it and see
            // the stale parity as "done": measured as wrong results on 4 % of the C2 queries) — and the two
            // sets then meet on a named barrier.
            // With two A groups per issuer the MMAs of my buffer may be up to TWO tiles behind my expansion:
            // both of my last two tiles are waited for, each on its own barrier (never more than one phase
            // outstanding per barrier, so the parity test stays unambiguous; a single barrier per buffer was
            // waited on one phase ahead here, read as "done", and the query tile was overwritten under
            // pending MMAs — wrong results for 64..512-bit signatures on large batches).
            if (t_idx > 0) {
                const uint32_t last = ((t_idx - 1u) & 1u) == part ? t_idx - 1u : t_idx - 2u;   // my set's last tile
                if (last < t_idx) {
                    if (QT == 2u) {   // that tile met both query tiles: buffers 0 and 1, use index = tile index
                        for (uint32_t b = 0; b < 2u; ++b) {
                            if (last >= 1u) mbar_wait(&acc_full[2u * b + ((last - 1u) & 1u)], ((last - 1u) >> 1) & 1u);
                            mbar_wait(&acc_full[2u * b + (last & 1u)], (last >> 1) & 1u);
                        }
                    } else {
                        const uint32_t ml = last >> 1;   // its index among the tiles of my buffer
                        if (ml >= 1u) mbar_wait(&acc_full[2u * part + ((ml - 1u) & 1u)], ((ml - 1u) >> 1) & 1u);
                        mbar_wait(&acc_full[2u * part + (ml & 1u)], (ml >> 1) & 1u);
                    }
                }
            }
            named_bar_sync(2, kMmaExpWarps * 32);
            for (uint32_t h = 0; h < QT; ++h) {
            const uint32_t q0 = (qg * QT + h) * NT + qoff;          // first query of MY part of query tile h
            for (uint32_t n = u; n < NB; n += kMmaTileRows) {
                const bool valid = q0 + n < p.Q;
                const uint64_t* src = p.queries + (size_t)(valid ? q0 + n : 0) * W;
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    const uint32_t nw = valid ? min(4u, W - 4u * kb) : 0u;
                    uint64_t w[4];
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) w[j] = j < nw ? __ldg(src + 4 * kb + j) : 0ull;
                    expand(base + L.b_off + (h * KB + kb) * NB * kMmaKBlockBytes, n, w, nw);
                }
            }
            if (!p.dense && part == 0) {
                for (uint32_t n = u; n < NB; n += kMmaTileRows) {
                    uint32_t e[8];
                    int T1, T2;
                    thr_split(__ldg(p.thrT + q0 + n), W, T1, T2);
                    encode_thr(T1, e);
                    const uint32_t dst = base + L.bthr_off + h * bthr_stride + (n >> 3) * 256u + (n & 7u) * 16u;
                    sts128(dst, e[0], e[1], e[2], e[3]);
                    sts128(dst + 128u, e[4], e[5], e[6], e[7]);
                    if (W > kMmaThrOneSlice) {   // second slice
                        encode_thr(T2, e);
                        sts128(dst + NB * 32u, e[0], e[1], e[2], e[3]);
                        sts128(dst + NB * 32u + 128u, e[4], e[5], e[6], e[7]);
                    }
                }
            }
            }   // query tiles of the item
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) arrive_in(b_full);
            // A tiles.  The expander warps work as TWO independent sets of kMmaExpWarps/2 warps, one per
            // MMA issuer / accumulator buffer (set s takes the tiles with t_idx % 2 == s, whose packed
            // rows always sit in stage s of the two-stage packed ring).  Expanding one tile is a latency
            // chain (barrier wait, shared loads, ~300 ALU ops, stores, proxy fence, arrive) that eight
            // warps together could not shorten; two tiles in flight hide half of it.
            for (uint32_t i = i0; i < i1; ++i, ++t_idx) {
                if ((t_idx & 1u) != part) continue;
                const uint32_t pst = part + 2u * pcur;      // my packed stage for this tile
                mbar_wait_relaxed(&p_full[pst], pph);
                if (u == 0 && part == 0) NPS_TRACE(1, 3);   // packed tile landed
                const uint64_t* my = reinterpret_cast<const uint64_t*>(smem + L.p_off + pst * L.p_stride) + (size_t)u * W;
                uint32_t lg = lg2[0];
                for (uint32_t g = 0; g < GPT; ++g, ++lg) {
                    const uint32_t slot = part * RG + lg % RG, par = (lg / RG) & 1u;
                    mbar_wait_relaxed(&g_empty[slot], par ^ 1u);
                    if (u == 0 && part == 0) NPS_TRACE(1, 1);   // group free
                    const uint32_t kb0 = g * GS, kb1 = min(kb0 + GS, KB);
                    for (uint32_t kb = kb0; kb < kb1; ++kb) {
                        // this thread's four words of the K-block: 16-byte loads when the row width is even
                        // (row strides of 64/128/256 bytes put every lane of a warp on the same banks; 8-byte
                        // loads cost ~1000 cycles per K-block there)
                        // rows past the end of the table expand whatever the stage holds: the epilogue
                        // ignores them
                        const uint32_t nw = min(4u, W - 4u * kb);      // uniform over the tile
                        uint64_t w[4] = {0ull, 0ull, 0ull, 0ull};
                        if ((W & 1u) == 0u) {
                            const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(my + 4 * kb);
                            w[0] = t.x;
                            w[1] = t.y;
                            if (nw > 2u) {
                                const ulonglong2 t2 = *reinterpret_cast<const ulonglong2*>(my + 4 * kb + 2);
                                w[2] = t2.x;
                                w[3] = t2.y;
                            }
                        } else {
#pragma unroll
                            for (uint32_t j = 0; j < 4; ++j)
                                if (j < nw) w[j] = my[4 * kb + j];
                        }
                        if (!(p.debug & 2u)) {
                            const uint32_t dst = base + L.a_off + (slot * GS + (kb - kb0)) * kMmaAStageBytes;
                            switch (nw) {
                                case 4: expand_row_words<4>(dst, u, w); break;
                                case 3: expand_row_words<3>(dst, u, w); break;
                                case 2: expand_row_words<2>(dst, u, w); break;
                                default: expand_row_words<1>(dst, u, w); break;
                            }
                        }
                    }
                    if (u == 0 && part == 0) NPS_TRACE(1, 4);   // group expanded (stores issued)
                    fence_proxy_async_smem();
                    if (u == 0 && part == 0) NPS_TRACE(1, 5);   // proxy fence done
                    __syncwarp();
                    if (lane == 0) arrive_in(&g_full[slot]);
                    if (u == 0 && part == 0) NPS_TRACE(1, 2);   // group written
                }
                lg2[0] = lg;
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_empty[pst]);
                if (++pcur == PH) {
                    pcur = 0;
                    pph ^= 1u;
                }
            }
        }
    } else {
        // =========================== epilogue: filter + append ====================================
        // 16 warps = 4 TMEM lane quarters (32 accumulator rows each) x 4 column parts (n_tile / 4
        // queries each).  Per tile and warp: ONE tcgen05.ld.x64 (measured: 16 warps x one x64 load
        // read a tile out in ~500 cycles, 8 warps x four pipelined x32 loads in ~900), 28 three-input
        // ANDs over the sign bits of  dot - T , one warp reduce.
        reg_inc<kRegsEpilogue>();
        const uint32_t ew = warp;
        const uint32_t quarter = warp & 3u;           // TMEM lanes this warp may read
        const uint32_t r_in_tile = quarter * 32u + lane;
        const uint32_t cpw = NT / 4u;                 // columns (queries) per warp: 8, 16, ..., 56
        const uint32_t c0 = (ew >> 2) * cpw;
        const uint32_t qmask = (1u << (cpw / 8u)) - 1u;   // 8-column quarters that exist
        uint2* hits = reinterpret_cast<uint2*>(smem + L.hitk_off) + ew * (kMmaHitsPerWarp + 1);
        float* thr_s = reinterpret_cast<float*>(smem + L.thr_off) + ew * 64u * QT;   // [QT][64]
        const uint32_t lane_lt = (1u << lane) - 1u;
        const int bits = (int)(W * 64u);
        uint32_t t_idx = 0;
        NPS_DEV(long long prof_wait = 0, prof_tile = 0, prof_slow = 0, prof_flush = 0; uint32_t prof_nslow = 0;
                const bool prof_on = p.trace != nullptr && blockIdx.x == 0;)
        uint32_t pend_q = 0xFFFFFFFFu, pend_pos = 0;   // append whose slot reservation is still in flight
        uint64_t pend_key = 0;
        for (uint32_t item = cta_id; item < total_items; item += num_ctas) {
            const uint32_t chunk = item / num_qgroups, qg = item - chunk * num_qgroups;
            const uint32_t i0 = chunk * p.tiles_per_chunk;
            const uint32_t i1 = min(i0 + p.tiles_per_chunk, ptiles);
            if constexpr (!kDense) {   // this warp's thresholds for the item's query tile(s)
                __syncwarp();
                for (uint32_t h = 0; h < QT; ++h)
                    for (uint32_t c = lane; c < cpw; c += 32u)
                        thr_s[h * 64u + c] = clamp_thr(__ldg(p.thrT + (qg * QT + h) * NT + c0 + c), W);
                __syncwarp();
            }
            for (uint32_t i = i0; i < i1; ++i, ++t_idx) {
                const uint32_t ti = kPair ? 2u * i + rank : i;      // this CTA's row tile of the phase
                const bool tile_ok = ti < p.phase_tiles;
                const unsigned long long row = (unsigned long long)phase_tile(p, tile_ok ? ti : 0u) * kMmaTileRows + r_in_tile;
                const bool valid = tile_ok && row < p.N;
                const uint64_t grow = p.row_base + row;
              for (uint32_t h = 0; h < QT; ++h) {   // the row tile against each query tile of the item
                // accumulator buffer and its use index: QT = 2 -> buffer h, every tile; else buffers alternate by tile
                const uint32_t buf = QT == 2u ? h : (t_idx & 1u);
                const uint32_t use = QT == 2u ? t_idx : (t_idx >> 1);
                const uint32_t q0 = (qg * QT + h) * NT + c0;          // first query of this warp's columns
                float* thr_h = thr_s + h * 64u;
                const uint32_t t_addr = tmem_base + ((quarter * 32u) << 16) + buf * NT + c0;
                uint32_t staged = 0;   // survivors of this tile waiting in the staging buffer
                NPS_DEV(const long long pc0 = prof_on ? clock64() : 0;)
                mbar_wait_relaxed(&acc_full[2u * buf + (use & 1u)], (use >> 1) & 1u);
                tc_fence_after();
                NPS_DEV(const long long pc1 = prof_on ? clock64() : 0; prof_wait += pc1 - pc0;)
                if (ew == 0 && lane == 0) NPS_TRACE(2, 1);   // accumulator ready

                uint32_t v[56];
                tmem_ld32_at<0>(t_addr, v);          // 56 columns = x32 + x16 + x8, one wait; columns beyond cpw
                tmem_ld16_at<32>(t_addr + 32u, v);   // belong to the neighbouring warp: read, then ignored
                tmem_ld8_at<48>(t_addr + 48u, v);
                tmem_ld_wait();
                if constexpr (kDense) {
                    // first phase: every distance of the tile goes to its fixed slot of the buffer
                    const size_t slot = (size_t)ti * kMmaTileRows + r_in_tile;
#pragma unroll
                    for (int j = 0; j < 56; ++j) {
                        if (tile_ok && (uint32_t)j < cpw && q0 + j < p.Q) {
                            const uint64_t d = (uint64_t)((bits - (int)__uint_as_float(v[j])) >> 1);
                            p.cand[(size_t)(q0 + j) * p.cap + slot] = valid ? ((d << kGlobalRowBits) | grow) : kNone;
                        }
                    }
                } else if (!(p.debug & 8u)) {
                    // bit i of m: 8-column quarter i holds a non-negative  dot - T  (sign bit clear)
                    uint32_t m = 0;
#pragma unroll
                    for (int qd = 0; qd < 7; ++qd) {
                        uint32_t acc = v[8 * qd] & v[8 * qd + 1];
#pragma unroll
                        for (int j = 2; j < 8; j += 2) acc &= v[8 * qd + j] & v[8 * qd + j + 1];
                        m |= (acc >> 31 ^ 1u) << qd;
                    }
                    m = (valid && !(p.debug & 1u)) ? (m & qmask) : 0u;
                    const uint32_t wm = __reduce_or_sync(0xffffffffu, m);   // quarters with a survivor in any lane
                    if (wm) {
                        NPS_DEV(const long long s0 = prof_on ? clock64() : 0;)
                        uint32_t wn = 0;   // survivors staged for this tile (warp-uniform)
                        if (wm & 1u) stage_hits8<0>(v, (m & 1u) != 0u, lane, wn, hits, lane_lt);
                        if (wm & 2u) stage_hits8<8>(v, (m & 2u) != 0u, lane, wn, hits, lane_lt);
                        if (wm & 4u) stage_hits8<16>(v, (m & 4u) != 0u, lane, wn, hits, lane_lt);
                        if (wm & 8u) stage_hits8<24>(v, (m & 8u) != 0u, lane, wn, hits, lane_lt);
                        if (wm & 16u) stage_hits8<32>(v, (m & 16u) != 0u, lane, wn, hits, lane_lt);
                        if (wm & 32u) stage_hits8<40>(v, (m & 32u) != 0u, lane, wn, hits, lane_lt);
                        if (wm & 64u) stage_hits8<48>(v, (m & 64u) != 0u, lane, wn, hits, lane_lt);
                        // staging overflow (loose early-phase thresholds): redo this warp's columns exactly
                        if (wn > (uint32_t)kMmaHitsPerWarp) {
                            rescan_group(t_addr, thr_h, q0, (int)min(cpw, 32u), grow, bits, valid, p.cnt, p.cand, p.cap);
                            if (cpw > 32u)
                                rescan_group(t_addr + 32u, thr_h + 32, q0 + 32u, (int)(cpw - 32u), grow, bits, valid, p.cnt,
                                             p.cand, p.cap);
                            wn = 0;
                        }
                        staged = wn;
                        NPS_DEV(if (prof_on) {
                            prof_slow += clock64() - s0;
                            ++prof_nslow;
                        })
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kPair) mbar_arrive_leader(&acc_empty[buf]);
                    else mbar_arrive(&acc_empty[buf]);
                }
                NPS_DEV(const long long pc2 = prof_on ? clock64() : 0; prof_tile += pc2 - pc1;)
                if (ew == 0 && lane == 0) NPS_TRACE(2, 2);   // accumulator released
                // Flush this warp's staging buffer, one entry per lane.  The atomic that reserves the
                // slot is issued now, the dependent store one tile later (pend_*), so the ~1 us
                // round trip of the returning atomic is never waited for.
                if constexpr (!kDense) {
                    NPS_DEV(const long long f0 = prof_on ? clock64() : 0;)
                    if (pend_q != 0xFFFFFFFFu) {
                        if (pend_pos < p.cap) p.cand[(size_t)pend_q * p.cap + pend_pos] = pend_key;
                        pend_q = 0xFFFFFFFFu;
                    }
                    const uint32_t n = staged;   // <= kMmaHitsPerWarp
                    if (n) {
                        __syncwarp();
                        // decode a staged entry: (dot - T, lane | column << 8) -> (query, key)
                        auto decode = [&](uint32_t e, uint32_t& q, uint64_t& key) {
                            const uint2 h = hits[e];
                            const uint32_t col = h.y >> 8;   // column within this warp's part
                            const int dot = (int)(__uint_as_float(h.x) + thr_h[col]);
                            const uint64_t r = grow - lane + (h.y & 31u);
                            q = q0 + col;
                            key = ((uint64_t)((bits - dot) >> 1) << kGlobalRowBits) | r;
                        };
                        for (uint32_t e = lane + 32u; e < n; e += 32u) {   // rare: more than 32 staged entries
                            uint32_t q;
                            uint64_t key;
                            decode(e, q, key);
                            global_append(p, q, key);
                        }
                        if (lane < n) {
                            decode(lane, pend_q, pend_key);
                            pend_pos = atomicAdd(p.cnt + pend_q, 1u);
                        }
                        __syncwarp();
                    }
                    NPS_DEV(if (prof_on) prof_flush += clock64() - f0;)
                }
              }   // query tiles of the item
            }
        }
        if (pend_q != 0xFFFFFFFFu && pend_pos < p.cap) p.cand[(size_t)pend_q * p.cap + pend_pos] = pend_key;
        NPS_DEV(if (prof_on && lane == 0 && ew < 8) {   // role slot 3 is shared with issuer 1: use its tail
            unsigned long long* o = p.trace + (3u * 4096u + 4000u + ew * 8u) * 2u;
            o[0] = (unsigned long long)prof_wait;
            o[1] = (unsigned long long)prof_tile;
            o[2] = (unsigned long long)prof_slow;
            o[3] = (unsigned long long)prof_flush;
            o[4] = prof_nslow;
            o[5] = t_idx;
        })
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();   // the peer's MMAs / remote arrivals no longer touch this CTA
    if (warp == kWarpMma) {
        if constexpr (kPair) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// Launch with (pdl = true) or without the programmatic-stream-serialization attribute.
template <typename Kern, typename Params>
static cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                              const Params& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_mma_scan(const MmaScanParams& p, int grid, size_t smem, cudaStream_t stream, bool pdl) {
    auto kern = p.pair ? (p.dense ? mma_scan_kernel<true, true> : mma_scan_kernel<false, true>)
                       : (p.dense ? mma_scan_kernel<true, false> : mma_scan_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (!p.pair) {
        e = launch_pdl(kern, dim3(grid), dim3(kMmaThreads), smem, stream, pdl, p);
    } else {   // clusters of two CTAs (one CTA pair per TPC); grid = 2 x pairs
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kMmaThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 2u : 1u;
        e = cudaLaunchKernelEx(&cfg, kern, p);
    }
    count_launch();
    return e;
}

static int g_pair_mode = 1;   // 0: never; 1: auto (two query tiles per item where they fit, else pairs where the single-CTA
                              // plan is squeezed by shared memory); 2: always; 3: always, one query tile per item
void mma_set_pair_mode(int mode) { g_pair_mode = mode; }

// =============================================================================================
// mma_select_kernel — fold a phase's candidates into the running exact top-k of every query.
// One CTA per query: keys of (previous best ++ candidates) are bitonic-sorted in shared memory
// (sort size = next power of two of the actual count), the k smallest become the new best,
// and the k-th one becomes the next phase's threshold.
// =============================================================================================
constexpr int kSelectThreads = 128;

__global__ void __launch_bounds__(kSelectThreads) mma_select_kernel(const MmaSelectParams p) {
    extern __shared__ __align__(16) uint64_t skeys[];
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t q = blockIdx.x;
    const int tid = threadIdx.x;
    if (q >= p.Q) {   // padding entries of the threshold arrays: never pass
        if (tid == 0) {
            p.thrT[q] = __int_as_float(0x7f800000);
            p.thrK[q] = 0ull;
        }
        return;
    }
    uint32_t cnt = p.fixed_cnt >= 0 ? (uint32_t)p.fixed_cnt : p.cnt[q];
    if (cnt > p.cap) {
        if (tid == 0) atomicAdd(p.overflow, 1u);
        cnt = p.cap;
    }
    const uint32_t nbest = p.first ? 0u : p.k;
    const uint32_t n = nbest + cnt;
    uint32_t S = 32;
    while (S < n) S <<= 1;
    for (uint32_t i = tid; i < S; i += kSelectThreads) {
        uint64_t key = kNone;
        if (i < nbest) key = p.best[(size_t)q * p.k + i];
        else if (i < n) key = p.cand[(size_t)q * p.cap + (i - nbest)];
        skeys[i] = key;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= S; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < (S >> 1); i += kSelectThreads) {
                const uint32_t pos = 2 * i - (i & (stride - 1));
                const uint64_t a = skeys[pos], b = skeys[pos + stride];
                const bool up = (pos & size) == 0;
                if ((a > b) == up) {
                    skeys[pos] = b;
                    skeys[pos + stride] = a;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = tid; i < p.k; i += kSelectThreads) {
        const uint64_t key = i < S ? skeys[i] : kNone;
        p.best[(size_t)q * p.k + i] = key;
        if (p.out_keys) p.out_keys[(size_t)q * p.k + i] = key;
        if (p.out_rows) p.out_rows[(size_t)q * p.k + i] = key == kNone ? kNone : (key & ((1ull << kGlobalRowBits) - 1ull));
        if (p.out_dists) p.out_dists[(size_t)q * p.k + i] = key == kNone ? kNone : (key >> kGlobalRowBits);
    }
    if (tid == 0) {
        const uint64_t kth = (p.k - 1 < S) ? skeys[p.k - 1] : kNone;
        p.thrK[q] = kth;
        // pass iff dist <= dist_k  <=>  dot >= bits - 2 dist_k   (no k-th key yet: everything passes)
        p.thrT[q] = kth == kNone ? __int_as_float(0xff800000)
                                 : (float)((int)p.bits - 2 * (int)(kth >> kGlobalRowBits));
        if (p.fixed_cnt < 0) p.cnt[q] = 0;
    }
}

// Small-k variant (k <= 32, the reference's top-10): one WARP per query, k rounds of "global minimum":
// every lane tracks the minimum of its strided slice of the keys in shared memory, a warp min-reduce
// picks the winner, the owning lane strikes it out and rescans its slice.  O(n + k * n / 32) instead of
// a full sort; 8 queries per CTA instead of one.
constexpr int kSelectWarps = 8;

__global__ void __launch_bounds__(kSelectWarps * 32) mma_select_warp_kernel(const MmaSelectParams p) {
    extern __shared__ __align__(16) uint64_t skeys[];
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kSelectWarps + warp;
    if (q >= p.Qpad) return;
    if (q >= p.Q) {   // padding entries of the threshold arrays: never pass
        if (lane == 0) {
            p.thrT[q] = __int_as_float(0x7f800000);
            p.thrK[q] = 0ull;
        }
        return;
    }
    uint64_t* keys = skeys + (size_t)warp * (p.k + p.cap);
    uint32_t cnt = p.fixed_cnt >= 0 ? (uint32_t)p.fixed_cnt : p.cnt[q];
    if (cnt > p.cap) {
        if (lane == 0) atomicAdd(p.overflow, 1u);
        cnt = p.cap;
    }
    const uint32_t nbest = p.first ? 0u : p.k;
    const uint32_t n = nbest + cnt;
    uint64_t my_min = kNone;
    uint32_t my_pos = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint64_t key = i < nbest ? p.best[(size_t)q * p.k + i] : p.cand[(size_t)q * p.cap + (i - nbest)];
        keys[i] = key;
        if (key < my_min) {
            my_min = key;
            my_pos = i;
        }
    }
    __syncwarp();
    uint64_t kth = kNone;
    for (uint32_t t = 0; t < p.k; ++t) {
        uint64_t m = my_min;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, m, o);
            m = other < m ? other : m;
        }
        if (m != kNone && my_min == m) {   // keys are unique: exactly one owner
            keys[my_pos] = kNone;
            my_min = kNone;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint64_t key = keys[i];
                if (key < my_min) {
                    my_min = key;
                    my_pos = i;
                }
            }
        }
        if (lane == 0) {
            const size_t o = (size_t)q * p.k + t;
            p.best[o] = m;
            if (p.out_keys) p.out_keys[o] = m;
            if (p.out_rows) p.out_rows[o] = m == kNone ? kNone : (m & ((1ull << kGlobalRowBits) - 1ull));
            if (p.out_dists) p.out_dists[o] = m == kNone ? kNone : (m >> kGlobalRowBits);
        }
        kth = m;
    }
    if (lane == 0) {
        p.thrK[q] = kth;
        p.thrT[q] = kth == kNone ? __int_as_float(0xff800000)
                                 : (float)((int)p.bits - 2 * (int)(kth >> kGlobalRowBits));
        if (p.fixed_cnt < 0) p.cnt[q] = 0;
    }
}

cudaError_t launch_mma_select(const MmaSelectParams& p, cudaStream_t stream, bool pdl) {
    if (p.k <= 32 && (size_t)(p.k + p.cap) * 8 * kSelectWarps <= 160 * 1024) {
        const size_t smem = (size_t)(p.k + p.cap) * 8 * kSelectWarps;
        cudaError_t e =
            cudaFuncSetAttribute(mma_select_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = launch_pdl(mma_select_warp_kernel, dim3((p.Qpad + kSelectWarps - 1) / kSelectWarps), dim3(kSelectWarps * 32),
                       smem, stream, pdl, p);
        count_launch();
        return e;
    }
    const size_t smem = (size_t)p.sort_cap * 8;
    cudaError_t e = cudaFuncSetAttribute(mma_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(mma_select_kernel, dim3(p.Qpad), dim3(kSelectThreads), smem, stream, pdl, p);
    count_launch();
    return e;
}


// Development knobs (schedule sweeps, traces, PDL on/off) are environment variables that exist only
// in builds made with NPS_KNOBS=1 or NPS_DEV=1 (np-sims_b200/build.py); a production build never
// reads the environment.
static inline const char* dev_knob(const char* name) {
#if defined(NPS_MMA_KNOBS) || defined(NPS_MMA_DEV)
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

// =============================================================================================
// host: plan + phase loop
// =============================================================================================
static uint32_t pow2_ceil(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

bool mma_plan(unsigned long long N, uint32_t W, uint32_t Q, uint32_t k, int sm_count, int smem_optin, MmaPlan* pl) {
    if (W == 0 || W > kMmaMaxWords || k == 0 || k > 1024 || Q == 0 || N == 0) return false;
    if (N > (1ull << 38)) return false;
    if (sm_count <= 0) sm_count = 148;
    if (smem_optin <= 0) smem_optin = 232448;
    // ---- operand tile: the widest query tile (multiple of 32, <= 256) that fits next to >= 2 A stages
    const uint32_t want = ((Q < kMmaMaxNTile ? Q : kMmaMaxNTile) + 31u) & ~31u;
    // A ring: each of the two MMA issuers owns RG groups of GS K-block stages; a group is filled,
    // consumed and released as a unit (one barrier wait + one tcgen05.commit).  Prefer whole-tile
    // groups; for wide signatures use smaller groups rather than a narrow query tile.
    const uint32_t KB = mma_kblocks(W);
    uint32_t n_tile = 0, gs_best = 0, rg_best = 0;
    auto select_tile = [&](uint32_t pair_mode) {
        n_tile = gs_best = rg_best = 0;
        for (uint32_t gs = KB;; gs = (gs + 1) / 2) {
            for (uint32_t nt = want; nt >= 32; nt -= 32) {
                if (nt < 128 && nt < want && gs > 1) break;          // try smaller groups before narrow tiles
                uint32_t rg = 0;
                for (uint32_t cand = 2; cand >= 1; --cand)
                    if (mma_smem_layout(W, nt, gs * 2 * cand, 2, pair_mode).total <= (uint32_t)smem_optin) {
                        rg = cand;
                        break;
                    }
                if (rg) {
                    n_tile = nt;
                    gs_best = gs;
                    rg_best = rg;
                    break;
                }
            }
            if (n_tile || gs == 1) break;
        }
    };
    // CTA pairs (cta_group::2): each CTA of a pair holds HALF of the query tile, which halves the B operand's
    // shared-memory footprint and read traffic.  Measured (profiles/r02_pair_vs_single.txt): a win exactly where
    // the single-CTA plan is squeezed by shared memory — a query tile narrower than asked for, or A groups
    // smaller than a row tile (1024-bit signatures +6 %, 2048-bit +46 %) — and neutral to slightly slower where
    // everything already fits (640-bit: -3 %, 256-bit: -7 %: the pair adds cross-CTA hand-offs to a kernel
    // that is bound by instruction issue in the expander and epilogue warps, not by B traffic).  So: auto.
    // Needs an even SM count (a pair shares a TPC) and at least two row tiles.
    select_tile(0);
    if (!n_tile) return false;
    const bool squeezed = n_tile < want || gs_best < KB;
    const bool can_pair = (sm_count % 2) == 0 && N > kMmaTileRows;
    uint32_t pair = can_pair && (g_pair_mode >= 2 || (g_pair_mode == 1 && squeezed)) ? 1u : 0u;
    if (const char* e = dev_knob("NPS_MMA_PAIR")) pair = atoi(e) != 0 && can_pair ? 1u : 0u;   // development
    if (pair) select_tile(1);
    // Two query tiles per work item (pairs only): each CTA keeps its halves of BOTH tiles resident and every
    // expanded row tile is multiplied with both, which halves the expansions per MMA — the expander warps'
    // instructions and shared-memory stores are what bounds the single-tile kernel.  Needs the plan with
    // whole-tile A groups, one per expander set, to still fit with the doubled B.
    uint32_t qt = 1;
    const char* pair_knob = dev_knob("NPS_MMA_PAIR");
    const bool pair_off = g_pair_mode == 0 || (pair_knob != nullptr && atoi(pair_knob) == 0);
    // Measured (profiles/r02_pair_vs_single.txt): exact, but NOT faster — 640-bit C2 last phase 0.608 ms against
    // 0.566 ms single-CTA: the step is paced by the hand-off chain expand -> issue -> retire per A slot and by
    // MMAs that run ~30 % slower while expander stores share the shared-memory port, not by the number of
    // expansions.  Kept behind mode 2 (measurements, tests); the automatic plan does not use it.
    const bool qt2_wanted = g_pair_mode == 2 || (pair_knob != nullptr && atoi(pair_knob) != 0);
    if (!pair_off && qt2_wanted && can_pair && Q > want && dev_knob("NPS_MMA_QT1") == nullptr) {
        for (uint32_t nt2 = want; nt2 >= 128 && nt2 + 32 >= want; nt2 -= 32)   // at most one step narrower
            if (mma_smem_layout(W, nt2, KB * 2, 2, 1, 2).total <= (uint32_t)smem_optin && Q > nt2) {
                pair = 1;
                qt = 2;
                n_tile = nt2;
                gs_best = KB;
                rg_best = 1;
                break;
            }
    }
    pl->pair = pair;
    pl->qt = qt;
    if (!n_tile) return false;
    if (const char* e = dev_knob("NPS_MMA_GROUPS")) {     // development: "GS,RG"
        uint32_t gs = 0, rg = 0;
        if (qt == 1 && sscanf(e, "%u,%u", &gs, &rg) == 2 && gs && rg >= 1 && rg <= 4) {
            gs_best = gs;
            rg_best = rg;
        }
    }
    pl->n_tile = n_tile;
    pl->group_kb = gs_best;
    pl->ring_per_issuer = rg_best;
    pl->a_stages = gs_best * 2 * rg_best;
    // packed-row ring: two stages (one per expander set) hide nothing of the bulk-copy latency — a set's next
    // tile is requested only when it has finished expanding the current one; take up to 8 if they fit
    uint32_t ps = 2;
    while (ps < 8 && mma_smem_layout(W, n_tile, pl->a_stages, ps + 2, pair, qt).total <= (uint32_t)smem_optin) ps += 2;
    if (const char* e = dev_knob("NPS_MMA_PSTAGES")) {   // development
        const uint32_t v = (uint32_t)atoi(e);
        if (v >= 2 && v % 2 == 0 && mma_smem_layout(W, n_tile, pl->a_stages, v, pair, qt).total <= (uint32_t)smem_optin) ps = v;
    }
    pl->p_stages = ps;
    pl->smem = mma_smem_layout(W, n_tile, pl->a_stages, ps, pair, qt).total;
    pl->num_qtiles = (Q + n_tile - 1) / n_tile;
    const uint32_t num_qgroups = (pl->num_qtiles + qt - 1) / qt;     // work items per row chunk
    pl->q_pad = num_qgroups * qt * n_tile;                           // an odd last group is padded with an empty tile
    // ---- candidate capacity and phase growth: expected survivors per phase ~ k * growth <= cap / 4
    uint32_t cap = pow2_ceil(16u * k);
    if (cap < 512u) cap = 512u;
    if (cap > 8192u) cap = 8192u;
    uint32_t growth = cap / (4u * k);
    if (growth > 8u) growth = 8u;
    if (growth < 2u) growth = 2u;
    // Small batches (a 10 000-query batch split over 8 GPUs leaves 6 query tiles per rank): every phase costs
    // ~25 us of launch + prologue + select whatever its size, and the epilogue has slack, so fewer, coarser
    // phases win — twice the candidate capacity (a denser first sample) and two doubling phases at the end
    // instead of three (measured, 400k x 640 bit, profiles/r02_phase_sweep.txt: 1250 queries 0.396 -> 0.354 ms,
    // 2500 queries 0.576 -> 0.535 ms; growth 16 is 3 % faster still but lets twice as much of the unsampled
    // tail of a sorted table through; from 10 000 queries on the finer schedule is the faster one).
    const bool small_batch = pl->num_qtiles <= 12 && k <= 32;
    if (small_batch) cap = 1024u;
    if (const char* e = dev_knob("NPS_MMA_CAP")) cap = (uint32_t)atoi(e);          // development
    if (const char* e = dev_knob("NPS_MMA_GROWTH")) growth = (uint32_t)atoi(e);    // development
    pl->cap = cap;
    pl->growth = growth;
    pl->sort_cap = pow2_ceil(k + cap);
    // ---- phases: nested tile subsets {t : t mod 2^e == 0}, e decreasing to 0.  The first (dense)
    // phase must fit the candidate buffer; early phases grow by `growth` (cheap, but their loose
    // thresholds let ~k * growth rows per query through), the last three only double, which keeps
    // the share of accumulator groups with a survivor (the epilogue's slow path) near 10 %.
    const unsigned long long T = (N + kMmaTileRows - 1) / kMmaTileRows;
    if (T > 0x7FFFFFFFull) return false;
    const uint32_t tiles0_max = cap / kMmaTileRows;
    uint32_t e0 = 0;
    while (((T + (1ull << e0) - 1) >> e0) > tiles0_max) ++e0;
    uint32_t gshift = 0;
    while ((2u << gshift) <= growth) ++gshift;        // growth rounded down to a power of two
    if (gshift < 1) gshift = 1;
    uint32_t exps[kMmaMaxPhases];
    uint32_t np = 0;
    exps[np++] = e0;
    for (uint32_t e = e0; e > 0;) {
        uint32_t late = small_batch ? 2u : 3u;   // number of final phases that only double
        if (const char* ev = dev_knob("NPS_MMA_LATE")) late = (uint32_t)atoi(ev);   // development
        const uint32_t step = e <= late ? 1u : (e - late < gshift ? e - late : gshift);
        e -= step;
        if (np == kMmaMaxPhases) return false;
        exps[np++] = e;
    }
    pl->num_phases = np;
    for (uint32_t ph = 0; ph < np; ++ph) {
        MmaPhase& P = pl->phase[ph];
        P.stride = 1u << exps[ph];
        const uint32_t J = (uint32_t)((T + P.stride - 1) / P.stride);
        if (ph == 0) {
            P.skip_mod = 0;
            P.tiles = J;
        } else {
            const uint32_t g = 1u << (exps[ph - 1] - exps[ph]);
            P.skip_mod = g;
            P.tiles = J - (J + g - 1) / g;
        }
        // chunks: balance waves of (chunk, query tile) items over the SMs (CTA pairs: over the SM
        // pairs, in units of two row tiles); ~1.5 units of fixed cost per item (query-tile expansion +
        // pipeline drain)
        const uint32_t units = pair ? (P.tiles + 1u) / 2u : P.tiles;
        const uint32_t workers = pair ? (uint32_t)sm_count / 2u : (uint32_t)sm_count;
        double best_cost = 1e300;
        uint32_t best_tpc = units ? units : 1;
        const uint32_t cmax = units < 4096u ? units : 4096u;
        for (uint32_t c = 1; c <= cmax; ++c) {
            const uint32_t tpc = (units + c - 1) / c;
            const uint32_t chunks = (units + tpc - 1) / tpc;
            const unsigned long long items = (unsigned long long)chunks * num_qgroups;
            const unsigned long long waves = (items + workers - 1) / workers;
            const double cost = (double)waves * ((double)tpc + 1.5);
            if (cost < best_cost) {
                best_cost = cost;
                best_tpc = tpc;
            }
        }
        P.tiles_per_chunk = best_tpc;
        P.num_chunks = units ? (units + best_tpc - 1) / best_tpc : 0;
        const unsigned long long items = (unsigned long long)P.num_chunks * num_qgroups;
        const uint32_t busy = (uint32_t)(items < (unsigned long long)workers ? items : (unsigned long long)workers);
        P.grid = pair ? 2u * busy : busy;
    }
    // ---- workspace
    size_t off = 0;
    pl->off_flag = off;
    off += 256;
    pl->off_cnt = off;
    off += align256((size_t)Q * 4);
    pl->off_thrT = off;
    off += align256((size_t)pl->q_pad * 4);
    pl->off_thrK = off;
    off += align256((size_t)pl->q_pad * 8);
    pl->off_best = off;
    off += align256((size_t)Q * k * 8);
    pl->off_cand = off;
    off += align256((size_t)Q * cap * 8);
    pl->total = off;
    return true;
}

cudaError_t run_mma_top_n(const MmaPlan& pl, const uint64_t* d_hashes, unsigned long long N, uint32_t W,
                          const uint64_t* d_queries, uint32_t Q, uint32_t k, unsigned long long row_base,
                          uint64_t* d_rows, uint64_t* d_dists, uint64_t* d_keys, void* ws, cudaStream_t stream,
                          cudaEvent_t* phase_events) {
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    uint32_t* flag = reinterpret_cast<uint32_t*>(w8 + pl.off_flag);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(w8 + pl.off_cnt);
    float* thrT = reinterpret_cast<float*>(w8 + pl.off_thrT);
    uint64_t* thrK = reinterpret_cast<uint64_t*>(w8 + pl.off_thrK);
    uint64_t* best = reinterpret_cast<uint64_t*>(w8 + pl.off_best);
    uint64_t* cand = reinterpret_cast<uint64_t*>(w8 + pl.off_cand);
    // flag + counters are contiguous at the start of the workspace
    cudaError_t e = cudaMemsetAsync(w8, 0, pl.off_thrT, stream);
    if (e != cudaSuccess) return e;
    // Programmatic dependent launch between the back-to-back scan / select launches, unless per-phase
    // events are being recorded between them (profiling) or NPS_NO_PDL is set (development).
    // ... nor under stream capture (a CUDA graph already removes the launch gaps; programmatic edges are not
    // needed there).
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap_status) != cudaSuccess) cap_status = cudaStreamCaptureStatusNone;
    const bool use_pdl = phase_events == nullptr && dev_knob("NPS_NO_PDL") == nullptr &&
                         cap_status == cudaStreamCaptureStatusNone;
    for (uint32_t ph = 0; ph < pl.num_phases; ++ph) {
        const MmaPhase& P = pl.phase[ph];
        const bool last = ph + 1 == pl.num_phases;
        if (P.tiles) {
            MmaScanParams sp;
            sp.hashes = d_hashes;
            sp.queries = d_queries;
            sp.thrT = thrT;
            sp.thrK = thrK;
            sp.cand = cand;
            sp.cnt = cnt;
            sp.overflow = flag;
            sp.N = N;
            sp.row_base = row_base;
            sp.Q = Q;
            sp.words = W;
            sp.cap = pl.cap;
            sp.n_tile = pl.n_tile;
            sp.num_qtiles = pl.num_qtiles;
            sp.phase_tiles = P.tiles;
            sp.tile_stride = P.stride;
            sp.skip_mod = P.skip_mod;
            sp.tiles_per_chunk = P.tiles_per_chunk;
            sp.num_chunks = P.num_chunks;
            sp.dense = ph == 0;
            sp.a_stages = pl.a_stages;
            sp.p_stages = pl.p_stages;
            sp.group_kb = pl.group_kb;
            sp.ring_per_issuer = pl.ring_per_issuer;
            sp.pair = pl.pair;
            sp.qt_per_item = pl.qt;
            sp.trace = nullptr;
            const char* trace_path = dev_knob("NPS_MMA_TRACE");
            const char* trace_phase = dev_knob("NPS_MMA_TRACE_PHASE");
            if (trace_path && (trace_phase ? (uint32_t)atoi(trace_phase) == ph : last)) {
                cudaMalloc(&sp.trace, 4 * 4096 * 16);
                cudaMemsetAsync(sp.trace, 0, 4 * 4096 * 16, stream);
            }
            {
                const char* dbg = dev_knob("NPS_MMA_DEBUG");
                sp.debug = (dbg && !sp.dense) ? (uint32_t)atoi(dbg) : 0u;
            }
            if (phase_events) {
                e = cudaEventRecord(phase_events[2 * ph], stream);
                if (e != cudaSuccess) return e;
            }
            // the scan of phase >= 1 directly follows the select of the phase before: let its prologue
            // overlap that kernel's tail (not when events or trace copies sit between the launches)
            e = launch_mma_scan(sp, (int)P.grid, pl.smem, stream, use_pdl && ph > 0 && !sp.trace);
            if (e != cudaSuccess) return e;
            if (phase_events) {
                e = cudaEventRecord(phase_events[2 * ph + 1], stream);
                if (e != cudaSuccess) return e;
            }
            if (sp.trace) {
                static unsigned long long host_trace[4 * 4096 * 2];
                cudaStreamSynchronize(stream);
                cudaMemcpy(host_trace, sp.trace, sizeof(host_trace), cudaMemcpyDeviceToHost);
                cudaFree(sp.trace);
                if (FILE* f = fopen(trace_path, "wb")) {
                    fwrite(host_trace, 1, sizeof(host_trace), f);
                    fclose(f);
                }
            }
        }
        MmaSelectParams se;
        se.cand = cand;
        se.cnt = cnt;
        se.best = best;
        se.thrT = thrT;
        se.thrK = thrK;
        se.overflow = flag;
        se.out_rows = last ? d_rows : nullptr;
        se.out_dists = last ? d_dists : nullptr;
        se.out_keys = last ? d_keys : nullptr;
        se.Q = Q;
        se.Qpad = pl.q_pad;
        se.k = k;
        se.cap = pl.cap;
        se.bits = W * 64;
        se.fixed_cnt = ph == 0 ? (int)(P.tiles * kMmaTileRows) : -1;
        se.first = ph == 0;
        se.sort_cap = pl.sort_cap;
        e = launch_mma_select(se, stream, use_pdl && P.tiles != 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace nps
