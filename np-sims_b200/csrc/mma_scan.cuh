// mma_scan.cuh — K2T: batched Hamming scan on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Why tensor cores for an XOR+popcount loop.  For a BATCH of queries the scan
//       d[i, q] = sum_w popc(H[i, w] ^ Q[q, w])
// is a matrix product over {-1,+1}: with bit b encoded as (1 - 2b),  <h_i, q_j> = bits - 2 d.
// On sm_100a the integer route costs 2 POPC per 64-bit word-pair on a 16-lane/clk/SM pipe
// (profiles/r01_scan_topk_c2_popc.txt: XU pipe 56 % busy, DRAM 0.01 %), i.e. the batched scan is
// compute-bound, not HBM-bound, and sm_100a has no native 1-bit MMA.  The
// closest native tensor-core form is kind::mxf4: E2M1 operands holding exactly +-1.0, unit block
// scales, FP32 accumulation in TMEM.  Every partial sum is an integer of magnitude <= bits < 2^24,
// so the result is EXACT — bit-identical distances — at 16384 MAC/clk/SM instead of 8 word-pairs.
// The signature table stays bit-packed in HBM/L2 (no expansion of traffic): tiles are expanded to
// E2M1 on chip, straight into the swizzled shared-memory operand layout.
//
// Kernel structure (one persistent CTA per SM, 28 warps, warp-specialised):
//   warps 0-15  epilogue  : one tcgen05.ld per tile and warp, sign test of  dot - threshold , staging
//   warps 16-23 expanders : bit-packed words -> E2M1 +-1, K-major SWIZZLE_128B operand tiles
//                           (B = the query tile, once per work item; A = row tile, per group)
//   warp 24     producer  : 1-D TMA bulk copies of bit-packed row tiles (128 rows) into a ring
//   warps 25-26 MMA issue : tcgen05.mma M=128 x N=n_tile x K=64, accumulators double-buffered in TMEM
// Selection: the kernel does no top-k bookkeeping.  Each launch ("phase") filters against a
// fixed per-query threshold (the exact k-th best key over the rows of all previous phases) and
// appends survivors (key = dist << 40 | row) to a per-query candidate buffer; mma_select_kernel
// then folds the candidates into the exact running top-k and tightens the threshold.  Phases
// visit the row tiles of the table in nested strided subsets, early ones ~8x larger than
// everything before them, the last three 2x, so few accumulator groups hold a survivor.
#pragma once
#include "tc.cuh"

namespace nps {

#ifndef NPS_MMA_EXP_WARPS
#define NPS_MMA_EXP_WARPS 8
#endif
#ifndef NPS_MMA_EPI_WARPS
#define NPS_MMA_EPI_WARPS 16
#endif
constexpr int kMmaExpWarps = NPS_MMA_EXP_WARPS;   // 4 or 8: threads per operand row = kMmaExpWarps / 4
constexpr int kMmaEpiWarps = NPS_MMA_EPI_WARPS;   // 16: 4 TMEM lane quarters x 4 column parts
constexpr int kMmaThreads = (4 + kMmaExpWarps + kMmaEpiWarps) * 32;   // + producer + 2 MMA issuers + 1 idle (full warpgroup)
constexpr int kMmaTileRows = 128;         // UMMA M
constexpr int kMmaKBlockBytes = 128;      // one swizzle row = 256 E2M1 = 256 signature bits = 4 words
constexpr uint32_t kMmaSfCol = 448;       // TMEM columns [448, 512): block scale factors (all 1.0)
constexpr uint32_t kMmaMaxNTile = 224;    // 2 accumulators x 224 columns = 448
constexpr int kMmaAStageBytes = kMmaTileRows * kMmaKBlockBytes;   // 16 KB
constexpr int kMmaHitsPerWarp = 64;
constexpr uint32_t kMmaMaxWords = 40;      // 2560 bits (the widest signatures the reference benchmarks)
constexpr uint32_t kMmaThrOneSlice = 23;   // up to here ONE folded-threshold slice covers |threshold| <= 64 * words
constexpr int kMmaThrSlice = 1510;         // what one K = 64 threshold slice can spell (even; 64 * 6 * 4 = 1536 max)
constexpr int kMmaThrRange = 2 * kMmaThrSlice;   // thresholds are clamped to +-this; > 64 * 40 bits

struct MmaScanParams {
    const uint64_t* hashes;    // [N, W] bit-packed
    const uint64_t* queries;   // [Q, W] bit-packed
    const float* thrT;         // [Qpad]  pass iff dot >= thrT[q]   (-inf: everything, +inf: nothing)
    const uint64_t* thrK;      // [Qpad]  the current k-th key (written by the select).  NOT read by the scan: rows tied with
                               //         it at the k-th distance are appended again and dropped by the select — applying
                               //         key < thrK[q] at flush time was measured (+5 % on C2) and not kept
    uint64_t* cand;            // [Q, cap] candidate keys
    uint32_t* cnt;             // [Q] candidates appended this phase
    uint32_t* overflow;        // [1] bumped if a warp's staging buffer overflowed (results then not final)
    unsigned long long N, row_base;
    uint32_t Q, words, cap, n_tile, num_qtiles;
    uint32_t phase_tiles;      // row tiles in this phase
    uint32_t tile_stride;      // tile = j * tile_stride
    uint32_t skip_mod;         // G: j runs over integers with j % G != 0 (0: j = i, first phase)
    uint32_t tiles_per_chunk, num_chunks;
    uint32_t dense;            // first phase: write every key at cand[q, i*128 + r], no filter
    uint32_t a_stages, p_stages;
    uint32_t group_kb, ring_per_issuer;   // A ring: 2 x ring_per_issuer groups of group_kb K-block stages
    uint32_t pair;             // CTA pairs (cta_group::2): phase_tiles counts row tiles, chunks count tile PAIRS
    uint32_t qt_per_item;      // query tiles per work item (2: pairs only — every expanded row tile meets two query tiles)
    unsigned long long* trace; // development only: per-role event timestamps of CTA 0
    uint32_t debug;            // development only: 1 = no epilogue filter, 2 = no expansion, 4 = no MMA
};

struct MmaSmem {
    uint32_t b_off, a_off, athr_off, bthr_off, p_off, p_stride, hitk_off, hitq_off, thr_off, wcnt_off, bar_off, tmem_off,
        total;
};
__host__ __device__ inline uint32_t mma_kblocks(uint32_t words) { return (words + 3) / 4; }
__host__ __device__ inline MmaSmem mma_smem_layout(uint32_t words, uint32_t n_tile, uint32_t a_stages,
                                                   uint32_t p_stages, uint32_t pair = 0, uint32_t qt = 1) {
    MmaSmem s;
    uint32_t off = 0;
    s.b_off = off;
    const uint32_t b_rows = pair ? n_tile / 2 : n_tile;   // a CTA of a pair holds half of the query tile
    off += qt * mma_kblocks(words) * b_rows * kMmaKBlockBytes;   // qt query tiles resident per work item
    s.a_off = off;
    off += a_stages * kMmaAStageBytes;
    s.athr_off = off;   // threshold operand tiles (no-swizzle K-major, one K = 64 slice): A = 4.0 everywhere,
    off += kMmaTileRows * 32;
    s.bthr_off = off;   // B row n = -(threshold of query n) / 4 spelled in E2M1 elements; a second tile
    off += qt * b_rows * 32 * (words > kMmaThrOneSlice ? 2u : 1u);   // (second slice) for wide signatures
    s.p_off = off;
    s.p_stride = (kMmaTileRows * words * 8 + 127u) & ~127u;
    off += p_stages * s.p_stride;
    s.hitk_off = off;
    off += kMmaEpiWarps * (kMmaHitsPerWarp + 1) * 8;   // (accumulator bits, position code) + 1 overflow slot
    s.hitq_off = off;
    s.thr_off = off;   // per epilogue warp: thresholds of its column groups
    off += qt * kMmaEpiWarps * 64u * 4u;
    s.wcnt_off = off;
    off += kMmaEpiWarps * 4;
    off = (off + 7u) & ~7u;
    s.bar_off = off;
    off += (2 * p_stages + 2 * a_stages + 7) * 8;
    s.tmem_off = off;
    off += 8;
    s.total = off + 1024;   // slack for the manual 1024-byte alignment of the base
    return s;
}

struct MmaSelectParams {
    uint64_t* cand;          // [Q, cap]
    uint32_t* cnt;           // [Q]   (reset to 0)
    uint64_t* best;          // [Q, k] running top-k keys, ascending, NPS_NONE padded (in/out)
    float* thrT;             // [Qpad] (out)
    uint64_t* thrK;          // [Qpad] (out)
    uint32_t* overflow;      // [1] incremented once per query whose buffer overflowed
    uint64_t* out_rows;      // final phase: decoded ids / distances / keys (any may be null)
    uint64_t* out_dists;
    uint64_t* out_keys;
    uint32_t Q, Qpad, k, cap, bits;
    int fixed_cnt;           // >= 0: dense phase, every query has exactly this many candidates
    uint32_t first;          // first phase: `best` holds garbage, treat as empty
    uint32_t sort_cap;       // shared-memory sort capacity (power of two >= k + cap)
};

cudaError_t launch_mma_scan(const MmaScanParams& p, int grid, size_t smem, cudaStream_t stream, bool pdl = false);
// development / measurement: 0 = single-CTA MMAs, 1 = CTA pairs where the shape allows (default)
void mma_set_pair_mode(int mode);
cudaError_t launch_mma_select(const MmaSelectParams& p, cudaStream_t stream, bool pdl = false);

// ---- host-side plan: tile shape, candidate capacity, phase schedule -------------------------
constexpr int kMmaMaxPhases = 24;
struct MmaPhase {
    uint32_t tiles, stride, skip_mod, tiles_per_chunk, num_chunks, grid;
};
struct MmaPlan {
    uint32_t n_tile, num_qtiles, q_pad, cap, growth, a_stages, p_stages, group_kb, ring_per_issuer, smem, sort_cap, num_phases;
    uint32_t pair;             // 1: CTA pairs (cta_group::2, M = 256); chunks of a phase then count tile pairs
    uint32_t qt;               // query tiles per work item (1 or 2)
    MmaPhase phase[kMmaMaxPhases];
    // workspace carve-up (bytes from the workspace base)
    size_t off_flag, off_cnt, off_thrT, off_thrK, off_best, off_cand, total;
};
// false if the tensor-core path does not apply to this shape (the caller uses the popc scan)
bool mma_plan(unsigned long long N, uint32_t W, uint32_t Q, uint32_t k, int sm_count, int smem_optin, MmaPlan* pl);
// d_rows/d_dists/d_keys: any may be null.  ws: pl.total bytes, 256-byte aligned.  ws[0] (uint32)
// counts the queries whose candidate buffer overflowed (results of those queries are not final).
cudaError_t run_mma_top_n(const MmaPlan& pl, const uint64_t* d_hashes, unsigned long long N, uint32_t W,
                          const uint64_t* d_queries, uint32_t Q, uint32_t k, unsigned long long row_base,
                          uint64_t* d_rows, uint64_t* d_dists, uint64_t* d_keys, void* ws, cudaStream_t stream,
                          cudaEvent_t* phase_events);   // nullable: 2 events per phase, recorded around each scan launch

}  // namespace nps
