// project_tc.cu — K1T: LSH hashing  sign(X . W^T) -> packed uint64  on the tensor cores
// (replaces lsh.index / lsh.index_one, np_sims/lsh.py:31-44, for float32 vectors).
//
// The projection is a dense contraction, so it runs as a TMA-fed tcgen05 GEMM: 128-row tiles of X
// and n_tile-projection tiles of W (both K-major fp32, SWIZZLE_128B boxes of 32 floats) stream
// through a shared-memory ring; one elected thread issues tcgen05.mma kind::tf32 (M128 x N x K8)
// into double-buffered TMEM accumulators; four epilogue warps read the accumulators back
// (thread = row), turn 64 signs into one uint64 word of the signature (bit j of the row =
// projection j >= 0, word j/64, bit j%64 — lsh.py:33) and store it.
//
// Exactness.  The reference computes the dot products in float64; TF32 keeps 10 mantissa bits of
// each operand, so an accumulator is only trusted when it is clearly away from zero:
//      |acc| >= eps * |x_i| * max_j |w_j|,   eps = 2^-9 + dims * 2^-22
// (truncation of both operands: <= 2^-10 relative each, summed against Cauchy-Schwarz; plus the
// fp32 accumulation).  Every other (row, projection) pair — a few percent — goes to a fix-up list
// and is recomputed in float64 by project_fix_kernel, which sets the bit the way the reference
// would.  Signatures therefore agree with the float64 path bit for bit (up to float64 summation
// order at |x.w| ~ 1e-16, as for project_f64.cu); the flipped-bit rate BEFORE fix-up is what the
// tests report as the TF32 tolerance.
#include <cuda.h>

#include <cstdio>
#include <mutex>

#include "project.cuh"
#include "tc.cuh"

namespace nps {

constexpr int kPtThreads = 6 * 32;          // producer, MMA issuer, 4 epilogue warps
constexpr int kPtTileRows = 128;
constexpr int kPtKBlock = 32;               // floats per K-block = one 128-byte swizzle row
constexpr int kPtAStage = kPtTileRows * 128;

struct ProjTcParams {
    uint64_t* out;              // [N, B/64]
    const float* norms;         // [N] |x_i|
    unsigned long long* fix;    // [fix_cap] row << 16 | projection
    unsigned int* fix_count;    // [1]
    unsigned long long N;
    uint32_t D, B, n_tile, num_ntiles, stages, fix_cap;
    float tol_scale;            // eps * max_j |w_j|
    unsigned long long num_mtiles;
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

__global__ void __launch_bounds__(kPtThreads, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const ProjTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const uint32_t NT = p.n_tile, S = p.stages;
    const uint32_t stage_bytes = kPtAStage + NT * 128u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = full + S;
    uint64_t* acc_full = empty + S;     // [2]
    uint64_t* acc_empty = acc_full + 2; // [2]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
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
        // ===== TMA producer: X tile (128 rows x 32 floats) + W tile (NT projections x 32 floats) per stage
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x) {
                const unsigned long long mt = item / p.num_ntiles;
                const uint32_t nt = (uint32_t)(item - mt * p.num_ntiles);
                for (uint32_t kb = 0; kb < KB; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    const uint32_t dst = base + s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], stage_bytes);
                    tma_load_2d(dst, &tm_x, kb * kPtKBlock, (uint32_t)(mt * kPtTileRows), &full[s]);
                    tma_load_2d(dst + kPtAStage, &tm_w, kb * kPtKBlock, nt * NT, &full[s]);
                    if (++s == S) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        const uint32_t idesc = umma_idesc(2, 2, kPtTileRows, NT);   // TF32 x TF32 -> F32
        uint32_t s = 0, ph = 0, t_idx = 0;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x, ++t_idx) {
            const uint32_t buf = t_idx & 1u;
            mbar_wait(&acc_empty[buf], ((t_idx >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * NT;
            for (uint32_t kb = 0; kb < KB; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (pt_elect_one()) {
                    const uint64_t a_desc = umma_desc_k_sw128(base + s * stage_bytes);
                    const uint64_t b_desc = umma_desc_k_sw128(base + s * stage_bytes + kPtAStage);
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j)    // K = 8 floats (32 bytes) per instruction
                        umma_tf32(d_tmem, a_desc + 2u * j, b_desc + 2u * j, idesc, (kb | j) != 0u);
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
    } else {
        // ===== epilogue: signs -> uint64 words, uncertain entries -> fix-up list
        const uint32_t quarter = warp & 3u;
        const uint32_t r_in_tile = quarter * 32u + lane;
        const uint32_t words_per_row = p.B / 64u;
        uint32_t t_idx = 0;
        for (unsigned long long item = blockIdx.x; item < total; item += gridDim.x, ++t_idx) {
            const unsigned long long mt = item / p.num_ntiles;
            const uint32_t nt = (uint32_t)(item - mt * p.num_ntiles);
            const uint32_t buf = t_idx & 1u;
            const unsigned long long row = mt * kPtTileRows + r_in_tile;
            const bool valid = row < p.N;
            const float tol = valid ? p.tol_scale * __ldg(p.norms + row) : 0.f;
            mbar_wait(&acc_full[buf], (t_idx >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((quarter * 32u) << 16) + buf * NT;
            for (uint32_t c0 = 0; c0 < NT; c0 += 64u) {
                uint32_t lohi[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + c0 + h * 32u, v);
                    tmem_ld_wait();
                    uint32_t bits = 0, unsure = 0;      // per column: sign, |acc| below the tolerance
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float a = __uint_as_float(v[j]);
                        bits |= (a >= 0.f ? 1u : 0u) << j;
                        unsure |= (fabsf(a) < tol ? 1u : 0u) << j;
                    }
                    lohi[h] = bits;
                    // Fix-up list: ONE global atomic per warp and 32-column group (a per-entry atomic on
                    // the single counter serialised in L2: measured 5 ms for 7 M entries).
                    const uint32_t mine = __popc(unsure);
                    uint32_t incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += up;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total) {   // warp-uniform
                        uint32_t basepos = 0;
                        if (lane == 31) basepos = atomicAdd(p.fix_count, total);
                        basepos = __shfl_sync(0xffffffffu, basepos, 31);
                        uint32_t pos = basepos + incl - mine;
                        const uint32_t col0 = nt * NT + c0 + h * 32u;
                        for (uint32_t m = unsure; m; m &= m - 1u, ++pos)
                            if (pos < p.fix_cap) p.fix[pos] = (row << 16) | (unsigned long long)(col0 + (__ffs(m) - 1));
                    }
                }
                if (valid)
                    p.out[row * words_per_row + (nt * NT + c0) / 64u] = (uint64_t)lohi[0] | ((uint64_t)lohi[1] << 32);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
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
        double s = 0.0;
        for (uint32_t k = lane; k < D; k += 32) s = fma((double)x[k], w[k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
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

bool project_tc_applicable(unsigned long long N, uint32_t D, uint32_t B) {
    return N > 0 && D >= 4 && D % 4 == 0 && B % 64 == 0 && B >= 64 && B <= 65536 && N < (1ull << 47) &&
           encode_tiled_fn() != nullptr;
}

size_t project_tc_workspace_bytes(unsigned long long N, uint32_t B, uint32_t* fix_cap_out) {
    // norms + counter + fix-up list sized for 1/8 of all (row, projection) pairs (a few % expected)
    unsigned long long cap = N * B / 8 + 1024;
    if (cap > 0x7FFFFFF0ull) cap = 0x7FFFFFF0ull;
    if (fix_cap_out) *fix_cap_out = (uint32_t)cap;
    return ((size_t)N * 4 + 255) / 256 * 256 + 256 + (size_t)cap * 8;
}

cudaError_t launch_project_tc(const float* X, unsigned long long N, uint32_t D, const float* W32, const double* W64,
                              uint32_t B, float w_norm_max, uint64_t* out, void* ws, int sm_count, int smem_optin,
                              cudaStream_t stream) {
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
    p.n_tile = (B % 256 == 0) ? 256 : (B % 128 == 0) ? 128 : 64;
    p.num_ntiles = B / p.n_tile;
    p.num_mtiles = (N + kPtTileRows - 1) / kPtTileRows;
    p.fix_cap = fix_cap;
    p.tol_scale = (1.0f / 512.0f + (float)D / 4194304.0f) * w_norm_max;
    const uint32_t stage_bytes = kPtAStage + p.n_tile * 128u;
    uint32_t stages = (uint32_t)((smem_optin - 2048) / (int)stage_bytes);
    if (stages > 6) stages = 6;
    if (stages < 2) return cudaErrorInvalidValue;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;

    CUtensorMap tm_x, tm_w;
    if (!make_tmap(&tm_x, X, N, D, kPtTileRows) || !make_tmap(&tm_w, W32, B, D, p.n_tile)) return cudaErrorInvalidValue;

    cudaError_t e = cudaMemsetAsync(fix_count, 0, 4, stream);
    if (e != cudaSuccess) return e;
    row_norm_kernel<<<(unsigned)((N + 7) / 8), 256, 0, stream>>>(X, N, D, norms);
    count_launch();
    e = cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned long long items = p.num_mtiles * p.num_ntiles;
    const int grid = (int)(items < (unsigned long long)sm_count ? items : (unsigned long long)sm_count);
    project_tc_kernel<<<grid, kPtThreads, smem, stream>>>(tm_x, tm_w, p);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    project_fix_kernel<<<sm_count * 8, 256, 0, stream>>>(fix, fix_count, fix_cap, X, W64, D, B,
                                                          reinterpret_cast<unsigned long long*>(out));
    count_launch();
    return cudaGetLastError();
}

}  // namespace nps
