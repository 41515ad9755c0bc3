// project_tc.cu — K1T: LSH hashing  sign(X . W^T) -> packed uint64  on the tensor cores
// (replaces lsh.index / lsh.index_one, np_sims/lsh.py:31-44, for float32 vectors).
//
// The projection is a dense contraction, so it runs as a TMA-fed tcgen05 GEMM: 128-row tiles of X
// and n_tile-projection tiles of W (K-major fp32, SWIZZLE_128B boxes of 32 floats) stream through
// a shared-memory ring; one elected thread issues tcgen05.mma kind::tf32 (M128 x N x K8) into
// double-buffered TMEM accumulators; four epilogue warps read the accumulators back (thread =
// row), turn 64 signs into one uint64 word of the signature (bit j of the row = projection j >= 0,
// word j/64, bit j%64 — lsh.py:33) and store it.
//
// Exactness.  The reference computes the dot products in float64.  A single TF32 product keeps 11
// significant bits per operand (relative error 2^-9 on the dot product: 2.8 % of all pairs at
// D = 300 would be too close to zero to trust).  So both operands are split into two TF32 terms
//      x = xh + xl (+ <= 2^-21 |x|)         w = wh + wl (+ <= 2^-22 |w|)
// and three products are accumulated:
//      C1 += xh.wh            C2 += xh.wl + xl.wh          result = C1 + C2  (epilogue)
// xh is what the tensor core makes of the raw float32 tile by itself — kind::tf32 ignores the low
// 13 mantissa bits (truncation; measured, scripts/ubench/umma_tf32_numerics.cu T1) — and four
// splitter warps write xl = rn_tf32(x - trunc(x)) beside each X tile in shared memory; W is split
// once per projection matrix from its float64 values (nps_project_split_projections).  The small
// correction terms have their own accumulator so that their roundings are relative to
// 2^-9 |x||w|, not |x||w|.  Accumulation model (same probe, T2-T5): inside one instruction the 8
// products and the accumulator are aligned to the largest exponent keeping 2 guard bits below
// float32's 24 (each addend truncated at 2^-25 of the largest), summed, and written back rounded
// toward zero: <= (9/4 + 1) * 2^-23 = 3.25 * 2^-23 per instruction, relative to the largest partial
// sum <= sum_i |x_i w_i| <= |x|_2 |w|_2.  Error budget relative to |x|_2 |w|_2:
//      representation  xl.wl + ex.w + x.ew      2^-21 + 2^-21 + 2^-22   =  5 * 2^-22
//      C1 accumulation D/8 instructions x 3.25 * 2^-23                   =  0.203 * D * 2^-22
//      C2 accumulation (2 D/8 instructions at 2^-9 of the scale), final add  <= (0.001 D + 0.25) * 2^-22
// An accumulator is trusted when |C1 + C2| >= eps |x_i| max_j |w_j|,  eps = (8 + D/4) * 2^-22
// (2.0e-5 at D = 300: 0.03 % of all pairs; 4.8e-5 at D = 768: 0.1 %).  Every other (row,
// projection) pair — and every row whose norm is not a comfortable float32 (zero, < 1e-18,
// > 1e18, inf, NaN) — is recomputed in float64 and its bit set the way the reference would: by the
// CTA's own fixer warps while the tile's rows are still in L2 (BF16 kernel: the usual handful per warp
// and item, through a shared-memory ring), or through a global fix-up list and project_fix_kernel.  If the list overflows (degenerate input), a
// gated launch of the float64 kernel redoes the whole call on the device.  Signatures therefore
// agree with the float64 path bit for bit (up to float64 summation order at |x.w| ~ 1e-16, as for
// project_f64.cu); the flipped-bit rate BEFORE fix-up is what the tests report as the tolerance.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "project.cuh"
#include "tc.cuh"

namespace nps {

constexpr int kPtThreads = 10 * 32;         // producer, MMA issuer, 4 epilogue warps, 4 splitter warps
constexpr int kPtTileRows = 128;
constexpr int kPtKBlock = 32;               // floats per K-block = one 128-byte swizzle row
constexpr int kPtAStage = kPtTileRows * 128;   // one X tile (xh or xl): 16 KB

struct ProjTcParams {
    uint64_t* out;              // [N, B/64]
    const float* norms;         // [N] |x_i|
    unsigned long long* fix;    // [fix_cap] row << 16 | projection
    unsigned int* fix_count;    // [2]: pairs sent to the list (saturating), pairs recomputed by the fixer warps
    unsigned long long N;
    uint32_t D, B, num_ntiles, stages, fix_cap;
    float tol_scale;            // eps * max_j |w_j|: the part of the tolerance that scales with |x_i|
    float acc_scale;            // BF16 kernel: factor of the accumulation term, sum_i |x_i[0 : 16 i)| qtab[i]
    const float* qtab;          // BF16 kernel: [4 * K-blocks] max_j |w_j[0 : 16 (i + 1))|
    unsigned long long num_mtiles;
    const float* X;             // [N, D] the rows again, for the float64 recomputation in the epilogue
    const double* W64;          // [B, D]
};

// float64 x_row . w_col by one warp, every lane gets the sum.  D % 4 == 0 on this path: 16-byte loads,
// four elements per lane and step, the loads of a row issued together.
__device__ __forceinline__ double fix_dot(const float* __restrict__ x, const double* __restrict__ w, uint32_t D,
                                          int lane) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
    for (uint32_t k = lane * 4u; k < D; k += 128u) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + k));
        const double2 wa = __ldg(reinterpret_cast<const double2*>(w + k));
        const double2 wb = __ldg(reinterpret_cast<const double2*>(w + k + 2));
        s0 = fma((double)xv.x, wa.x, s0);
        s1 = fma((double)xv.y, wa.y, s1);
        s0 = fma((double)xv.z, wb.x, s0);
        s1 = fma((double)xv.w, wb.y, s1);
    }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
// Unsure entries on their way from the epilogue warps to the fixer warps of the same CTA: one
// single-producer single-consumer ring per epilogue warp (quarter).  The producer never blocks: what does not
// fit goes to the global fix-up list instead.
constexpr uint32_t kPtFixQ = 128;
struct FixQueue {
    uint32_t head[4], tail[4], done[4], pad[4];
    unsigned long long ent[4][kPtFixQ];   // row << 16 | projection
};

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t c0, uint32_t c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ bool pt_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// xl = rn_tf32(x - xh), xh = x with the low 13 mantissa bits cleared (what kind::tf32 reads from x):
// x - xh - xl <= 2^-21 |x|
__device__ __forceinline__ uint32_t tf32_low_part(uint32_t x) {
    return rn_tf32(__uint_as_float(x) - __uint_as_float(x & 0xFFFFE000u));
}

// Epilogue shared by both operand encodings: signs of C1 + C2 -> uint64 words; uncertain entries go to
// the CTA's fixer warps (fq, BF16 kernel) or to the global fix-up list (one atomic per warp and item).  Accumulator buffer b holds C1 at columns [2 NT b, +NT)
// and C2 right behind it.
// norm_s: null -> |x_i| comes from p.norms (row_norm_kernel); else the splitter warps left it in shared
// memory, kPtNormBufs buffers of 128 floats indexed by the item count (BF16 kernel).
constexpr uint32_t kPtNormBufs = 8;   // splitters run at most 2 + stages items ahead of the epilogue
template <uint32_t NT, uint32_t NBUF = 2, bool kRelaxed = false>
__device__ __forceinline__ void project_epilogue(const ProjTcParams& p, uint32_t tmem_base, uint64_t* acc_full,
                                                 uint64_t* acc_empty, int warp, int lane, unsigned long long total,
                                                 const float* norm_s = nullptr, FixQueue* fq = nullptr,
                                                 const float* acc_s = nullptr) {
        const uint32_t quarter = warp & 3u;
        const uint32_t r_in_tile = quarter * 32u + lane;
        const uint32_t words_per_row = p.B / 64u;
        constexpr uint32_t G = NT / 32u;       // 32-column groups per item
        uint32_t t_idx = 0, q_head = 0;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x, ++t_idx) {
            const unsigned long long mt = item / p.num_ntiles;
            const uint32_t nt = (uint32_t)(item - mt * p.num_ntiles);
            const uint32_t buf = t_idx % NBUF;
            const unsigned long long row = mt * kPtTileRows + r_in_tile;
            const bool valid = row < p.N;
            float norm = (valid && norm_s == nullptr) ? __ldg(p.norms + row) : 1.f;
            if (kRelaxed) mbar_wait_relaxed(&acc_full[buf], (t_idx / NBUF) & 1u);   // sleeps in hardware, no polling
            else mbar_wait(&acc_full[buf], (t_idx / NBUF) & 1u);
            tc_fence_after();
            if (valid && norm_s != nullptr)      // written before the splitters' last split_done of the item
                norm = reinterpret_cast<const volatile float*>(norm_s)[(t_idx % kPtNormBufs) * kPtTileRows + r_in_tile];
            const bool tame = norm >= 1e-18f && norm <= 1e18f;     // false for 0, tiny, huge, inf, NaN
            // BF16 kernel: + the accumulation term the splitters summed up for this row (see the kernel's header)
            float tol = p.tol_scale * norm;
            if (acc_s != nullptr)
                tol = fmaf(p.acc_scale, reinterpret_cast<const volatile float*>(acc_s)[(t_idx % kPtNormBufs) * kPtTileRows + r_in_tile], tol);
            const uint32_t t_addr = tmem_base + ((quarter * 32u) << 16) + buf * (2u * NT);
            // Between acc_full and the release below the next item's MMAs wait (the BF16 kernel's accumulator is single-
            // buffered): three instructions per value here — the add, a funnel shift that collects the sign bit, a
            // running minimum of |a| — and the per-value tolerance test only for the 16-column groups whose minimum is
            // below the tolerance (a = -0.0 has the sign bit set but is below every tolerance, so it is recomputed).
            uint32_t sign_m[G], unsure_m[G];
#pragma unroll
            for (uint32_t g = 0; g < G; ++g) {
                uint32_t neg = 0, unsure = 0;      // neg: sign bits, first column in bit 31
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h) {  // 16 columns at a time: 32 live registers, not 64
                    uint32_t v1[16], v2[16];
                    tmem_ld16_at<0>(t_addr + g * 32u + h * 16u, v1);
                    tmem_ld16_at<0>(t_addr + NT + g * 32u + h * 16u, v2);
                    tmem_ld_wait();
                    float a[16], lowest = __int_as_float(0x7f800000);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        a[j] = __uint_as_float(v1[j]) + __uint_as_float(v2[j]);
                        neg = __funnelshift_l(__float_as_uint(a[j]), neg, 1);
                        lowest = fminf(lowest, fabsf(a[j]));
                    }
                    if (!(lowest >= tol)) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) unsure |= (fabsf(a[j]) < tol ? 1u : 0u) << (h * 16u + j);
                    }
                }
                sign_m[g] = ~__brev(neg);           // bit j = (a_j >= 0) wherever the accumulator is trusted
                unsure_m[g] = !valid ? 0u : (tame ? unsure : 0xFFFFFFFFu);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);   // accumulator free before the global traffic below
            uint32_t mine = 0;
#pragma unroll
            for (uint32_t g = 0; g < G; ++g) mine += __popc(unsure_m[g]);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            uint32_t all = __shfl_sync(0xffffffffu, incl, 31);
            if (valid) {
#pragma unroll
                for (uint32_t g = 0; g < G; g += 2)
                    p.out[row * words_per_row + (nt * NT + g * 32u) / 64u] =
                        (uint64_t)sign_m[g] | ((uint64_t)sign_m[g + 1] << 32);
            }
            if (all != 0u && fq != nullptr) {   // warp-uniform
                // The usual case (0.1 % of the pairs: a handful per warp and item): hand them to this quarter's fixer
                // warp, which recomputes them in float64 while the tile's rows are still in L2 and patches the
                // words stored above.  (A separate pass over a list re-read every such row from DRAM: 40 GB at C3.)
                const uint32_t tail = *reinterpret_cast<volatile uint32_t*>(&fq->tail[quarter]);
                if (q_head + all - tail <= kPtFixQ) {
                    uint32_t pos = q_head + incl - mine;
#pragma unroll
                    for (uint32_t g = 0; g < G; ++g)
                        for (uint32_t m = unsure_m[g]; m; m &= m - 1u, ++pos)
                            fq->ent[quarter][pos % kPtFixQ] =
                                (row << 16) | (unsigned long long)(nt * NT + g * 32u + (__ffs(m) - 1));
                    __threadfence_block();     // words and entries before the new head
                    __syncwarp();
                    q_head += all;
                    if (lane == 0) {
                        *reinterpret_cast<volatile uint32_t*>(&fq->head[quarter]) = q_head;
                        atomicAdd(p.fix_count + 1, all);     // statistics only
                    }
                    all = 0;
                }
            }
            if (all) {   // no fixer warps (TF32 kernel), ring full, degenerate rows: the list and project_fix_kernel
                // Saturating: once the list has overflowed (count > capacity: the gated float64 launch redoes the whole
                // call) nothing more is added, so the 32-bit counter cannot wrap on degenerate multi-million-row
                // input (every pair unsure) and hide the overflow.
                uint32_t basepos = p.fix_cap;
                if (lane == 31 && *reinterpret_cast<volatile unsigned int*>(p.fix_count) <= p.fix_cap)
                    basepos = atomicAdd(p.fix_count, all);
                basepos = __shfl_sync(0xffffffffu, basepos, 31);
                uint32_t pos = basepos + incl - mine;
#pragma unroll
                for (uint32_t g = 0; g < G; ++g)
                    for (uint32_t m = unsure_m[g]; m; m &= m - 1u, ++pos)
                        if (pos < p.fix_cap)
                            p.fix[pos] = (row << 16) | (unsigned long long)(nt * NT + g * 32u + (__ffs(m) - 1));
            }
        }
        if (fq != nullptr && lane == 0) *reinterpret_cast<volatile uint32_t*>(&fq->done[quarter]) = 1u;
}

// Fixer warp f of kFixers: drains rings f, f + kFixers, ... until their epilogue warps are done.
// Entries a fixer warp recomputes together.  C3 on one B200: no fixer warps (list + project_fix_kernel only) 68.4-69.8 ms,
// 1 -> 65.2-65.5 (28 % of the entries overflow the rings to the list), 2 -> 65.1-65.3 (none overflow), 4 -> 66.4-66.6.
constexpr uint32_t kPbFixBatch = 2;
template <uint32_t kFixers>
__device__ __forceinline__ void project_fixer(const ProjTcParams& p, FixQueue* fq, uint32_t f, int lane) {
    constexpr int kRings = 4 / kFixers;
    uint32_t tail[kRings] = {};
    uint32_t live = (1u << kRings) - 1u;
    while (live) {
        bool idle = true;
#pragma unroll
        for (int h = 0; h < kRings; ++h) {
            if (!(live >> h & 1u)) continue;
            const uint32_t q = f + kFixers * h;
            uint32_t done = 0, head = 0;
            if (lane == 0) {
                done = *reinterpret_cast<volatile uint32_t*>(&fq->done[q]);
                __threadfence_block();
                head = *reinterpret_cast<volatile uint32_t*>(&fq->head[q]);
            }
            done = __shfl_sync(0xffffffffu, done, 0);
            head = __shfl_sync(0xffffffffu, head, 0);
            if (head == tail[h]) {
                if (done) live &= ~(1u << h);
                continue;
            }
            idle = false;
            __threadfence_block();
            while (tail[h] != head) {
                // kPbFixBatch entries at a time: one warp on one entry is a chain of L2 round trips and does not keep
                // up with the epilogue; the loads of a batch overlap.  A short batch repeats its first entry (unused).
                const uint32_t n = min(head - tail[h], kPbFixBatch);
                unsigned long long code[kPbFixBatch];
                const float* x[kPbFixBatch];
                const double* w[kPbFixBatch];
#pragma unroll
                for (uint32_t e = 0; e < kPbFixBatch; ++e) {
                    code[e] = fq->ent[q][(tail[h] + (e < n ? e : 0u)) % kPtFixQ];
                    x[e] = p.X + (code[e] >> 16) * p.D;
                    w[e] = p.W64 + (size_t)(code[e] & 0xFFFFu) * p.D;
                }
                double s0[kPbFixBatch] = {}, s1[kPbFixBatch] = {};
#pragma unroll 2
                for (uint32_t k = lane * 4u; k < p.D; k += 128u) {
#pragma unroll
                    for (uint32_t e = 0; e < kPbFixBatch; ++e) {
                        const float4 xv = __ldg(reinterpret_cast<const float4*>(x[e] + k));
                        const double2 wa = __ldg(reinterpret_cast<const double2*>(w[e] + k));
                        const double2 wb = __ldg(reinterpret_cast<const double2*>(w[e] + k + 2));
                        s0[e] = fma((double)xv.x, wa.x, s0[e]);
                        s1[e] = fma((double)xv.y, wa.y, s1[e]);
                        s0[e] = fma((double)xv.z, wb.x, s0[e]);
                        s1[e] = fma((double)xv.w, wb.y, s1[e]);
                    }
                }
                double mine = 0.0;      // lane e ends up with the sum of entry e
#pragma unroll
                for (uint32_t e = 0; e < kPbFixBatch; ++e) {
                    double t = s0[e] + s1[e];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (lane == (int)e) mine = t;
                }
                unsigned long long mycode = code[0];
#pragma unroll
                for (uint32_t e = 1; e < kPbFixBatch; ++e)
                    if (lane == (int)e) mycode = code[e];
                if ((uint32_t)lane < n) {
                    const unsigned long long row = mycode >> 16;
                    const uint32_t col = (uint32_t)(mycode & 0xFFFFu);
                    unsigned long long* word =
                        reinterpret_cast<unsigned long long*>(p.out) + row * (p.B / 64u) + col / 64u;
                    const unsigned long long mask = 1ull << (col & 63u);
                    if (mine >= 0.0) atomicOr(word, mask);
                    else atomicAnd(word, ~mask);
                }
                tail[h] += n;
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&fq->tail[q]) = tail[h];
            }
        }
        if (idle) __nanosleep(500);
    }
}

// Stage layout: [xh tile 16 KB][xl tile 16 KB][wh tile NT x 128 B][wl tile NT x 128 B]
template <uint32_t NT>
__global__ void __launch_bounds__(kPtThreads, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const ProjTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t S = p.stages;
    constexpr uint32_t kWTile = NT * 128u;
    constexpr uint32_t stage_bytes = 2u * kPtAStage + 2u * kWTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
    uint64_t* full = bars;              // TMA landed (X raw + wh + wl)
    uint64_t* split_done = full + S;    // splitter warps wrote xl
    uint64_t* empty = split_done + S;   // MMAs of the stage retired
    uint64_t* acc_full = empty + S;     // [2]
    uint64_t* acc_empty = acc_full + 2; // [2]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split_done[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_s);

    const uint32_t KB = (p.D + kPtKBlock - 1) / kPtKBlock;
    const unsigned long long total = p.num_mtiles * p.num_ntiles;

    if (warp == 0) {
        // ===== TMA producer: X tile (128 rows x 32 floats) + the wh and wl tiles (NT projections x 32 floats)
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
                const unsigned long long mt = item / p.num_ntiles;
                const uint32_t nt = (uint32_t)(item - mt * p.num_ntiles);
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    const uint32_t dst = base + s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], kPtAStage + 2u * kWTile);
                    tma_load_2d(dst, &tm_x, kb * kPtKBlock, (uint32_t)(mt * kPtTileRows), &full[s]);
                    tma_load_2d(dst + 2u * kPtAStage, &tm_w, kb * kPtKBlock, nt * NT, &full[s]);
                    tma_load_2d(dst + 2u * kPtAStage + kWTile, &tm_w, kb * kPtKBlock, p.B + nt * NT, &full[s]);
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: C1 += xh.wh ; C2 += xh.wl + xl.wh
        // The wh and wl tiles are contiguous in the stage and so are C1 and C2 in TMEM: ONE instruction of
        // width 2 NT does xh.[wh; wl] -> [C1 | C2] and reads xh from shared memory once.
        const uint32_t idesc = umma_idesc(2, 2, kPtTileRows, NT);        // TF32 x TF32 -> F32, N = NT
        const uint32_t idesc2 = umma_idesc(2, 2, kPtTileRows, 2u * NT);  // N = 2 NT
        uint32_t s = 0, ph = 0, t_idx = 0;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x, ++t_idx) {
            const uint32_t buf = t_idx & 1u;
            mbar_wait(&acc_empty[buf], ((t_idx >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t c1 = tmem_base + buf * (2u * NT), c2 = c1 + NT;
            for (uint32_t kb = 0; kb < KB; ++kb) {
                mbar_wait(&full[s], ph);
                mbar_wait(&split_done[s], ph);
                tc_fence_after();
                if (pt_elect_one()) {
                    const uint32_t st = base + s * stage_bytes;
                    const uint64_t xh = umma_desc_k_sw128(st), xl = umma_desc_k_sw128(st + kPtAStage);
                    const uint64_t wh = umma_desc_k_sw128(st + 2u * kPtAStage);   // wl follows at + kWTile
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {   // K = 8 floats (32 bytes) per instruction
                        umma_tf32(c1, xh + 2u * j, wh + 2u * j, idesc2, (kb | j) != 0u);
                        umma_tf32(c2, xl + 2u * j, wh + 2u * j, idesc, true);
                    }
                    umma_commit(&empty[s]);
                    if (kb + 1 == KB) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp >= 6) {
        // ===== splitters: X tile (fp32 as TMA left it, = xh to the tensor core) -> xl beside it.
        // Elementwise, so the 128-byte swizzle of the tile is irrelevant; 16-byte chunks, conflict-free.
        const uint32_t sidx = (uint32_t)tid - 192u;
        uint32_t s = 0, ph = 0;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
            for (uint32_t kb = 0; kb < KB; ++kb) {
                mbar_wait(&full[s], ph);
                const uint4* xh = reinterpret_cast<const uint4*>(smem + s * stage_bytes);
                uint4* xl = reinterpret_cast<uint4*>(smem + s * stage_bytes) + kPtAStage / 16;
#pragma unroll
                for (uint32_t i = 0; i < kPtAStage / 16 / 128; ++i) {
                    const uint4 v = xh[sidx + i * 128u];
                    uint4 l;
                    l.x = tf32_low_part(v.x);
                    l.y = tf32_low_part(v.y);
                    l.z = tf32_low_part(v.z);
                    l.w = tf32_low_part(v.w);
                    xl[sidx + i * 128u] = l;
                }
                fence_proxy_async_smem();   // generic-proxy stores -> visible to tcgen05.mma (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&split_done[s]);
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        project_epilogue<NT>(p, tmem_base, acc_full, acc_empty, warp, lane, total);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// BF16 x 2 variant (the default): the same three-product scheme with both operands split into two
// BF16 terms, x = x0 + x1 (+ <= 2^-18 |x|), w = w0 + w1, on kind::f16 — twice the MMA rate of
// kind::tf32 and half the operand bytes per flop, which is what bounds the TF32 kernel (shared-
// memory bandwidth).  X still arrives as float32 by TMA (two 32-float boxes per 64-float K-block);
// the splitter warps (thread = row) convert it into the two K-major SWIZZLE_128B BF16 tiles.
// Accumulation model measured for kind::f16 too (scripts/ubench/umma_bf16_numerics.cu: 2 guard
// bits, RZ write-back): per instruction (K = 16) (17/4 + 1) * 2^-23.  Budget relative to |x||w|:
//      representation  x1.w1 + ex.w + x.ew        3 * 2^-18               =  48 * 2^-22
//      C1 accumulation D/16 instructions x 5.25 * 2^-23                   =  0.164 * D * 2^-22
//      C2 accumulation (at 2^-8 of the scale), final add                  <= (0.001 D + 0.25) * 2^-22
// The accumulation term is relative to the largest partial sum an instruction sees, and instruction i (elements
// [16 i, 16 i + 16)) sees at most sum_{k < 16 (i+1)} |x_k w_k| <= |x[0 : 16 (i+1))| |w[0 : 16 (i+1))|: the PREFIX
// norms (Cauchy-Schwarz on the prefix), not the full ones.  The splitter thread of a row has the running sum of
// squares anyway; it adds up  P = sum_i |x[0 : 16 (i+1))| qtab[i],  qtab[i] = max_j |w_j[0 : 16 (i+1))| (48 floats
// at D = 768, from split_projections), and the row's tolerance is
//      (52 + 0.006 D) * 2^-22 * |x| max_j |w_j|  +  2.625 * 1.02 * 2^-22 * P
// (1.02: |x0_k| <= (1 + 2^-8) |x_k|, the same for w0, float32 rounding of P).  For rows and projections without
// structure P ~ |x||w| D/32, half of the full-norm bound: 0.096 % of the pairs are unsure at D = 768 with the
// full norms, about 0.06 % with the prefix norms.
// Stage layout: [raw fp32 2 x 16 KB][x0 16 KB][x1 16 KB][w0 NT x 128 B][w1 NT x 128 B]
// =============================================================================================
constexpr int kPbKBlock = 64;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // lo -> bits [0,16), hi -> bits [16,32)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// two floats -> (x0 pair, x1 pair): x0 = bf16_rn(x), x1 = bf16_rn(x - x0)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& h, uint32_t& l) {
    h = pack_bf16x2(a, b);
    l = pack_bf16x2(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xFFFF0000u));
}

// ---------------------------------------------------------------------------------- BF16 x 2, X operand in TMEM
// Both operands split into two BF16 terms on kind::f16 (twice the MMA rate of kind::tf32, half the operand bytes
// per flop).  X arrives as float32 by TMA into a raw ring; four splitter warps (thread = row = TMEM lane)
// convert each 64-float K-block into x0 = rn_bf16(x), x1 = rn_bf16(x - x0) and
// write them STRAIGHT INTO TENSOR MEMORY (tcgen05.st, 32 + 32 packed columns); the MMAs take their A operand
// from there (tcgen05.mma [d], [a], b-desc).  The split X tiles never touch shared memory: per K-block it
// carries 32 KB raw X in + 32 KB splitter reads + 32 KB W in + 48 KB W reads by the MMAs = 144 KB (208 KB with
// the operands in shared memory, which was what bounded that version: C3 65 -> 50 ms).
// TMEM: one accumulator pair [C1 | C2] in columns [0, 2 NT) — the epilogue frees it after ~0.3 us of tcgen05.ld,
// so single buffering costs ~2 % of an item — and the A ring in columns 256 + 64 s.
// Warps: 0 X producer, 1 MMA issuer, 2-5 epilogue, 6-9 splitters, 10 W producer, 11-14 fixers (15 warps: four on
// a sub-partition, hence 128 registers).
// Tried and measured on C3 (10M x 768 -> 1024 bit; this layout: 50-52 ms): a second splitter group on alternate
// K-blocks with one fixer warp left (main kernel -10 %, but 11 M instead of 3 M pairs fall through to
// project_fix_kernel: 53.7 ms); 3 raw slots + 3 W stages (61.7 ms: the W ring needs its four stages); CTA pairs
// (cta_group::2, each CTA staging half of every W tile, 2-5 raw slots: 54-69 ms, the deeper the raw ring the slower).
constexpr uint32_t kTsRawSlots = 2;
constexpr uint32_t kTsACol = 256;
constexpr int kPbThreads = 15 * 32;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

template <uint32_t NT>
__global__ void __launch_bounds__(kPbThreads, 1)
project_bf16_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                    const ProjTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t S = p.stages;                                // W stages in shared memory = A slots in TMEM (<= 4)
    constexpr uint32_t R = kTsRawSlots;
    constexpr uint32_t kWTile = NT * 128u;                      // NT projections x 64 bf16
    constexpr uint32_t kRaw = 2u * kPtAStage;                   // 32 KB raw
    constexpr uint32_t w_bytes = 2u * kWTile;
    const uint32_t w_base = R * kRaw;                           // W ring behind the raw ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + w_base + S * w_bytes);
    uint64_t* raw_full = bars;              // [R] TMA landed the float32 tile
    uint64_t* raw_empty = raw_full + R;     // [R] splitters are done with it
    uint64_t* w_full = raw_empty + R;       // [S] TMA landed w0, w1
    uint64_t* split_done = w_full + S;      // [S] splitters wrote x0, x1 into TMEM slot s
    uint64_t* op_empty = split_done + S;    // [S] MMAs of the slot retired (W stage and A slot free)
    uint64_t* acc_full = op_empty + S;      // [1]
    uint64_t* acc_empty = acc_full + 1;     // [1]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_empty + 1);
    float* norm_s = reinterpret_cast<float*>(tmem_ptr_s + 4);   // [kPtNormBufs][128]: |x_i| from the splitters
    float* acc_s = norm_s + kPtNormBufs * kPtTileRows;          // [kPtNormBufs][128]: sum_i |x_i[0 : 16 i)| qtab[i]
    FixQueue* fq = reinterpret_cast<FixQueue*>(acc_s + kPtNormBufs * kPtTileRows);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 16) fq->head[tid] = 0u;   // head, tail, done, pad
    if (tid == 0) {
        for (uint32_t s = 0; s < R; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], 4);
        }
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&split_done[s], 4);
            mbar_init(&op_empty[s], 1);
        }
        mbar_init(&acc_full[0], 1);
        mbar_init(&acc_empty[0], 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_s);

    const uint32_t KB = (p.D + kPbKBlock - 1) / kPbKBlock;
    const unsigned long long total = p.num_mtiles * p.num_ntiles;

    if (warp == 0) {
        // ===== X producer: two 32-float boxes per K-block into the raw ring
        if (lane == 0) {
            uint32_t rs = 0, rph = 0;
            for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
                const unsigned long long mt = item / p.num_ntiles;
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    mbar_wait_relaxed(&raw_empty[rs], rph ^ 1u);
                    const uint32_t dst = base + rs * kRaw;
                    mbar_arrive_expect_tx(&raw_full[rs], kRaw);
                    tma_load_2d(dst, &tm_x, kb * kPbKBlock, (uint32_t)(mt * kPtTileRows), &raw_full[rs]);
                    tma_load_2d(dst + kPtAStage, &tm_x, kb * kPbKBlock + 32u, (uint32_t)(mt * kPtTileRows), &raw_full[rs]);
                    if (++rs == R) {
                        rs = 0;
                        rph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 10) {
        // ===== W producer: w0 and w1 tiles of the K-block
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
                const unsigned long long mt = item / p.num_ntiles;
                const uint32_t nt = (uint32_t)(item - mt * p.num_ntiles);
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    mbar_wait_relaxed(&op_empty[s], ph ^ 1u);
                    const uint32_t dst = base + w_base + s * w_bytes;
                    mbar_arrive_expect_tx(&w_full[s], w_bytes);
                    tma_load_2d(dst, &tm_w, kb * kPbKBlock, nt * NT, &w_full[s]);
                    tma_load_2d(dst + kWTile, &tm_w, kb * kPbKBlock, p.B + nt * NT, &w_full[s]);
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: [C1 | C2] += x0.[w0; w1] (one instruction of width 2 NT) ; C2 += x1.w0, A from TMEM
        const uint32_t idesc = umma_idesc(1, 1, kPtTileRows, NT);        // BF16 x BF16 -> F32
        const uint32_t idesc2 = umma_idesc(1, 1, kPtTileRows, 2u * NT);
        uint32_t s = 0, ph = 0, t_idx = 0;
        const uint32_t c1 = tmem_base, c2 = c1 + NT;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x, ++t_idx) {
            mbar_wait(&acc_empty[0], (t_idx & 1u) ^ 1u);
            tc_fence_after();
            for (uint32_t kb = 0; kb < KB; ++kb) {
                mbar_wait(&w_full[s], ph);
                mbar_wait(&split_done[s], ph);
                tc_fence_after();
                if (pt_elect_one()) {
                    const uint32_t a0 = tmem_base + kTsACol + s * 64u, a1 = a0 + 32u;
                    const uint64_t w0 = umma_desc_k_sw128(base + w_base + s * w_bytes);   // w1 follows at + kWTile
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) {   // K = 16 bf16 per instruction: 8 TMEM columns, 32 bytes of W
                        umma_f16_ts(c1, a0 + 8u * j, w0 + 2u * j, idesc2, (kb | j) != 0u);
                        umma_f16_ts(c2, a1 + 8u * j, w0 + 2u * j, idesc, true);
                    }
                    umma_commit(&op_empty[s]);
                    if (kb + 1 == KB) umma_commit(&acc_full[0]);
                }
                __syncwarp();
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp >= 11) {
        project_fixer<4>(p, fq, (uint32_t)warp - 11u, lane);
    } else if (warp >= 6) {
        // ===== splitters: thread = row = TMEM lane (a warp reaches the 32 lanes of its own quarter, warp % 4).
        // Shared-memory reads: the row's 128-byte line with the chunk index XORed by (row & 7), conflict-free for
        // the eight rows of a quarter warp.
        const uint32_t quarter = (uint32_t)warp & 3u;
        const uint32_t r = quarter * 32u + (uint32_t)lane, r7 = r & 7u;
        const uint32_t rowoff = (r >> 3) * 1024u + r7 * 128u;
        const uint32_t t_lane = tmem_base + ((quarter * 32u) << 16) + kTsACol;
        uint32_t s = 0, ph = 0, rs = 0, rph = 0, t_idx = 0;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x, ++t_idx) {
            float sumsq = 0.f;      // this row sees all of its D floats pass by: |x_i| for free (no norm kernel)
            float pacc = 0.f;       // sum over the MMA instructions (16 elements each) of |x_i[0 : 16 (i + 1))| qtab[i]
            for (uint32_t kb = 0; kb < KB; ++kb) {
                const float4 qt = __ldg(reinterpret_cast<const float4*>(p.qtab) + kb);   // in flight during the waits
                mbar_wait(&raw_full[rs], rph);
                mbar_wait(&op_empty[s], ph ^ 1u);          // the slot's previous MMAs retired
                tc_fence_after();
                const uint8_t* rawp = smem + rs * kRaw;
                const uint32_t t_slot = t_lane + s * 64u;
#pragma unroll
                for (uint32_t jj = 0; jj < 4; ++jj) {      // 16 floats -> 8 packed columns of x0 and of x1
                    uint32_t h[8], l[8];
#pragma unroll
                    for (uint32_t u = 0; u < 2; ++u) {
                        const uint32_t j = 2u * jj + u;
                        const uint8_t* box = rawp + (j >> 2) * kPtAStage + rowoff;
                        const uint32_t c0 = 2u * (j & 3u);
                        const float4 a = *reinterpret_cast<const float4*>(box + ((c0 ^ r7) << 4));
                        const float4 b = *reinterpret_cast<const float4*>(box + (((c0 + 1u) ^ r7) << 4));
                        split_bf16x2(a.x, a.y, h[4 * u + 0], l[4 * u + 0]);
                        split_bf16x2(a.z, a.w, h[4 * u + 1], l[4 * u + 1]);
                        split_bf16x2(b.x, b.y, h[4 * u + 2], l[4 * u + 2]);
                        split_bf16x2(b.z, b.w, h[4 * u + 3], l[4 * u + 3]);
                        sumsq = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, sumsq))));
                        sumsq = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, sumsq))));
                    }
                    tmem_st8(t_slot + 8u * jj, h);
                    tmem_st8(t_slot + 32u + 8u * jj, l);
                    float prefix;      // |x[0 : 16 (i + 1))|: MUFU.SQRT (2 ulp) is plenty against the 1.02
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(prefix) : "f"(sumsq));
                    pacc = fmaf(prefix, jj == 0 ? qt.x : jj == 1 ? qt.y : jj == 2 ? qt.z : qt.w, pacc);
                }
                if (kb + 1 == KB) {
                    norm_s[(t_idx % kPtNormBufs) * kPtTileRows + r] = sqrtf(sumsq) * 1.001f;
                    acc_s[(t_idx % kPtNormBufs) * kPtTileRows + r] = pacc;
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&raw_empty[rs]);
                    mbar_arrive(&split_done[s]);
                }
                if (++rs == R) {
                    rs = 0;
                    rph ^= 1u;
                }
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        project_epilogue<NT, 1, true>(p, tmem_base, acc_full, acc_empty, warp, lane, total, norm_s, fq, acc_s);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// W (float64) -> both operand encodings, back to back in one buffer (project_split_bytes):
//   [2, B, D]  float32  wh = rn_tf32(w), wl = rn_tf32(w - wh)          w - wh - wl <= 2^-22 |w|
//   [2, B, Dp] bfloat16 w0 = rn_bf16(w), w1 = rn_bf16(w - w0), Dp = D rounded up to 8 (16-byte rows for
//              the tensor map), zero padded                             w - w0 - w1 <= 2^-18 |w|
//   [4 * ceil(D / 64)] float32  qtab[i] = max_j |w_j[0 : 16 (i + 1))|_2: prefix norms for the BF16 kernel's tolerance
static size_t split_tf32_bytes(uint32_t B, uint32_t D) { return ((size_t)2 * B * D * 4 + 255) / 256 * 256; }
static uint32_t split_dp(uint32_t D) { return (D + 7u) & ~7u; }
static size_t split_bf16_bytes(uint32_t B, uint32_t D) { return ((size_t)2 * B * split_dp(D) * 2 + 255) / 256 * 256; }
static uint32_t split_qtab_len(uint32_t D) { return (D + 63u) / 64u * 4u; }     // one entry per MMA instruction
size_t project_split_bytes(uint32_t B, uint32_t D) {
    return split_tf32_bytes(B, D) + split_bf16_bytes(B, D) + (size_t)split_qtab_len(D) * 4;
}

__global__ void __launch_bounds__(256) split_projections_kernel(const double* __restrict__ W, uint32_t B, uint32_t D,
                                                                uint32_t Dp, float* __restrict__ out,
                                                                uint16_t* __restrict__ out16) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;    // over [B, Dp]
    if (i >= (size_t)B * Dp) return;
    const uint32_t row = (uint32_t)(i / Dp), col = (uint32_t)(i - (size_t)row * Dp);
    const double w = col < D ? W[(size_t)row * D + col] : 0.0;
    if (col < D) {
        const uint32_t hi = rn_tf32((float)w);
        const uint32_t lo = rn_tf32((float)(w - (double)__uint_as_float(hi)));
        out[(size_t)row * D + col] = __uint_as_float(hi);
        out[(size_t)B * D + (size_t)row * D + col] = __uint_as_float(lo);
    }
    const uint32_t b0 = pack_bf16x2((float)w, 0.f) & 0xFFFFu;
    const uint32_t b1 = pack_bf16x2((float)(w - (double)__uint_as_float(b0 << 16)), 0.f) & 0xFFFFu;
    out16[i] = (uint16_t)b0;
    out16[(size_t)B * Dp + i] = (uint16_t)b1;
}

// qtab[i] = max over the projections of |w_j[0 : 16 (i + 1))|_2 (non-negative floats order like their bit patterns):
// one warp per projection, lane l owns the 16-element blocks l, l + 32, ...
__global__ void __launch_bounds__(256) prefix_norm_kernel(const double* __restrict__ W, uint32_t B, uint32_t D,
                                                          uint32_t len, float* __restrict__ qtab) {
    const uint32_t j = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= B) return;
    double before = 0.0;        // sum of squares of all blocks before this round of 32
    for (uint32_t i0 = 0; i0 < len; i0 += 32) {
        const uint32_t i = i0 + lane;
        double mine = 0.0;
        for (uint32_t k = 16u * i; k < 16u * i + 16u && k < D; ++k) {
            const double w = W[(size_t)j * D + k];
            mine += w * w;
        }
        double incl = mine;     // inclusive prefix sum over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (i < len) atomicMax(reinterpret_cast<unsigned int*>(qtab + i), __float_as_uint((float)sqrt(before + incl) * 1.0001f));
        before += __shfl_sync(0xffffffffu, incl, 31);
    }
}

cudaError_t launch_split_projections(const double* W64, uint32_t B, uint32_t D, float* out, cudaStream_t stream) {
    const size_t n = (size_t)B * split_dp(D);
    if (n == 0) return cudaSuccess;
    uint16_t* out16 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(out) + split_tf32_bytes(B, D));
    split_projections_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(W64, B, D, split_dp(D), out, out16);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    float* qtab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(out16) + split_bf16_bytes(B, D));
    e = cudaMemsetAsync(qtab, 0, (size_t)split_qtab_len(D) * 4, stream);
    if (e != cudaSuccess) return e;
    prefix_norm_kernel<<<(B + 7) / 8, 256, 0, stream>>>(W64, B, D, split_qtab_len(D), qtab);
    count_launch();
    return cudaGetLastError();
}

// |x_i| for the tolerance test
__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ X, unsigned long long N, uint32_t D,
                                                       float* __restrict__ norms) {
    const unsigned long long row = (unsigned long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= N) return;
    const float* x = X + row * D;
    float s = 0.f;
    for (uint32_t k = threadIdx.x & 31; k < D; k += 32) s = fmaf(x[k], x[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) norms[row] = sqrtf(s) * 1.0001f;
}

// float64 recomputation of the listed (row, projection) pairs: one warp per entry
__global__ void __launch_bounds__(256) project_fix_kernel(const unsigned long long* __restrict__ fix,
                                                          const unsigned int* __restrict__ fix_count, uint32_t fix_cap,
                                                          const float* __restrict__ X, const double* __restrict__ W,
                                                          uint32_t D, uint32_t B, unsigned long long* __restrict__ out) {
    const unsigned int n = min(*fix_count, fix_cap);
    const int lane = threadIdx.x & 31;
    for (unsigned int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += gridDim.x * 8) {
        const unsigned long long code = fix[e];
        const unsigned long long row = code >> 16;
        const uint32_t col = (uint32_t)(code & 0xFFFFu);
        const float* x = X + row * D;
        const double* w = W + (size_t)col * D;
        const double s = fix_dot(x, w, D, lane);
        if (lane == 0) {
            unsigned long long* word = out + row * (B / 64u) + col / 64u;
            const unsigned long long mask = 1ull << (col & 63u);
            if (s >= 0.0) atomicOr(word, mask);
            else atomicAnd(word, ~mask);
        }
    }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// [rows, D] float32 row-major, boxes of 32 floats x box_rows rows, 128-byte swizzle, zero fill
static bool make_tmap(CUtensorMap* tm, const float* ptr, unsigned long long rows, uint32_t D, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {D, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)D * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kPtKBlock, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows, Dp] bfloat16 row-major, boxes of 64 elements x box_rows rows, 128-byte swizzle, zero fill
static bool make_tmap_bf16(CUtensorMap* tm, const void* ptr, unsigned long long rows, uint32_t Dp, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {Dp, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)Dp * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kPbKBlock, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool project_tc_applicable(unsigned long long N, uint32_t D, uint32_t B) {
    return N > 0 && D >= 4 && D % 4 == 0 && B % 64 == 0 && B >= 64 && B <= 65536 && N < (1ull << 47) &&
           encode_tiled_fn() != nullptr;
}

size_t project_tc_workspace_bytes(unsigned long long N, uint32_t B, uint32_t* fix_cap_out) {
    // norms + counter + fix-up list sized for 1/64 of all (row, projection) pairs: 0.03-0.4 % are expected
    // (DESIGN.md), rows with an unusable norm add B entries each; beyond that the gated float64 launch
    // takes over.  (The single-TF32 version needed 1/8: 10 GB for the C3 table.)
    unsigned long long cap = N * B / 64 + 4096;
    if (cap > 0x7FFFFFF0ull) cap = 0x7FFFFFF0ull;
    if (fix_cap_out) *fix_cap_out = (uint32_t)cap;
    return ((size_t)N * 4 + 255) / 256 * 256 + 256 + (size_t)cap * 8;
}

cudaError_t launch_project_tc(const float* X, unsigned long long N, uint32_t D, const float* Wsplit, const double* W64,
                              uint32_t B, float w_norm_max, uint64_t* out, void* ws, int sm_count, int smem_optin,
                              cudaStream_t stream, bool use_bf16) {
    uint32_t fix_cap = 0;
    project_tc_workspace_bytes(N, B, &fix_cap);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    float* norms = reinterpret_cast<float*>(w8);
    const size_t norms_bytes = ((size_t)N * 4 + 255) / 256 * 256;
    unsigned int* fix_count = reinterpret_cast<unsigned int*>(w8 + norms_bytes);
    unsigned long long* fix = reinterpret_cast<unsigned long long*>(w8 + norms_bytes + 256);

    ProjTcParams p;
    p.out = out;
    p.norms = norms;
    p.fix = fix;
    p.fix_count = fix_count;
    p.N = N;
    p.D = D;
    p.B = B;
    p.X = X;
    p.W64 = W64;
    const uint32_t n_tile = (B % 128 == 0) ? 128 : 64;
    p.num_ntiles = B / n_tile;
    p.num_mtiles = (N + kPtTileRows - 1) / kPtTileRows;
    p.fix_cap = fix_cap;
    // eps: (8 + D/4) * 2^-22 for the TF32 split; BF16: (52 + 0.006 D) * 2^-22 + the prefix-norm accumulation term (see the kernels)
    p.tol_scale = (use_bf16 ? 52.0f + 0.006f * (float)D : 8.0f + 0.25f * (float)D) / 4194304.0f * w_norm_max;
    p.acc_scale = w_norm_max < 1e-20f ? 0.f : 2.625f * 1.02f / 4194304.0f;   // (tiny norm: the raw-signs measurement mode)
    p.qtab = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(Wsplit) + split_tf32_bytes(B, D) + split_bf16_bytes(B, D));
    // TF32: stages of [xh | xl | wh | wl]; BF16: a raw-X ring of kTsRawSlots x 32 KB + W stages [w0 | w1] (X operand in TMEM)
    const uint32_t stage_bytes = (use_bf16 ? 0u : 2u * kPtAStage) + 2u * n_tile * 128u;
    const uint32_t fixed = use_bf16 ? kTsRawSlots * 2u * kPtAStage + 2u * kPtNormBufs * kPtTileRows * 4u + (uint32_t)sizeof(FixQueue) : 0u;
    uint32_t stages = (uint32_t)((smem_optin - 2048 - (int)fixed) / (int)stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) return cudaErrorInvalidValue;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + fixed + 1024 + 512;

    CUtensorMap tm_x, tm_w;   // tm_w: the split projections [2B, D] (high rows, then low rows)
    if (!make_tmap(&tm_x, X, N, D, kPtTileRows)) return cudaErrorInvalidValue;
    if (use_bf16) {
        const void* w16 = reinterpret_cast<const uint8_t*>(Wsplit) + split_tf32_bytes(B, D);
        if (!make_tmap_bf16(&tm_w, w16, 2ull * B, split_dp(D), n_tile)) return cudaErrorInvalidValue;
    } else if (!make_tmap(&tm_w, Wsplit, 2ull * B, D, n_tile)) {
        return cudaErrorInvalidValue;
    }

    cudaError_t e = cudaMemsetAsync(fix_count, 0, 8, stream);
    if (e != cudaSuccess) return e;
    if (!use_bf16) {   // the BF16 kernel's splitter warps compute |x_i| on the way
        row_norm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, stream>>>(X, N, D, norms);
        count_launch();
    }
    auto kern = use_bf16 ? (n_tile == 128 ? project_bf16_kernel<128> : project_bf16_kernel<64>)
                         : (n_tile == 128 ? project_tc_kernel<128> : project_tc_kernel<64>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned long long items = p.num_mtiles * p.num_ntiles;
    const int grid = (int)(items < (unsigned long long)sm_count ? items : (unsigned long long)sm_count);
    kern<<<grid, use_bf16 ? kPbThreads : kPtThreads, smem, stream>>>(tm_x, tm_w, p);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    project_fix_kernel<<<sm_count * 8, 256, 0, stream>>>(fix, fix_count, fix_cap, X, W64, D, B,
                                                          reinterpret_cast<unsigned long long*>(out));
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // fix-up list overflowed (degenerate input): redo the whole call in float64, on the device
    return launch_project_f64(X, 1, N, D, W64, B, out, nullptr, stream, fix_count, fix_cap);
}

}  // namespace nps
