// Probe: numerics of tcgen05.mma kind::tf32 on this GPU — what project_tc.cu's error budget assumes.
//   T1  are fp32 operand bits below TF32's 10 mantissa bits truncated or rounded?
//   T2  how many guard bits survive the alignment of the K = 8 products inside one instruction?
//   T3  the same across instructions (accumulator re-read from TMEM)
//   T4  rounding of the accumulator write-back
// D[m][0] = sum_k A[m][k] * 1 with hand-made rows of A; printed as (D - 1) in units of 2^-23.
#include <cmath>
#include <cstdio>
#include <vector>
#include "tc.cuh"
using namespace nps;

__global__ void __launch_bounds__(128, 1) k_probe(const float* A, const float* B, uint32_t N, float* D) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    auto put = [&](uint8_t* tile, const float* src, uint32_t rows) {   // row-major [rows][32 floats] -> SWIZZLE_128B
        for (uint32_t i = threadIdx.x; i < rows * 8; i += blockDim.x) {
            const uint32_t r = i >> 3, c = i & 7;
            *reinterpret_cast<uint4*>(tile + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) =
                *reinterpret_cast<const uint4*>(src + r * 32 + c * 4);
        }
    };
    put(smem, A, 128);
    put(smem + 16384, B, N);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(2, 2, 128, N);
        const uint64_t da = umma_desc_k_sw128(base), db = umma_desc_k_sw128(base + 16384);
        for (int j = 0; j < 4; ++j) umma_tf32(tmem, da + 2 * j, db + 2 * j, idesc, j != 0);
        umma_commit(&bar); mbar_wait(&bar, 0);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((warp * 32u) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)threadIdx.x * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    const uint32_t N = 32;
    std::vector<float> A(128 * 32, 0.f), B(N * 32, 0.f);
    for (int k = 0; k < 32; ++k) B[0 * 32 + k] = 1.f;
    const float lowbits = 1.f + ldexpf(1.f, -11) + ldexpf(1.f, -13);
    B[1 * 32 + 0] = lowbits;
    A[0 * 32 + 0] = lowbits;                                 // T1
    A[2 * 32 + 0] = -lowbits;
    A[3 * 32 + 0] = 1.f + ldexpf(1.f, -11) - ldexpf(1.f, -20);   // just below half an ulp
    for (int i = 0; i < 10; ++i) {                           // T2: 1 + 7 * 2^-(22+i) inside one instruction
        A[(8 + i) * 32 + 0] = 1.f;
        for (int k = 1; k < 8; ++k) A[(8 + i) * 32 + k] = ldexpf(1.f, -(22 + i));
    }
    for (int i = 0; i < 10; ++i) {                           // T3: 1 (instr 0) + 8 * 2^-(22+i) (instr 1)
        A[(24 + i) * 32 + 0] = 1.f;
        for (int k = 8; k < 16; ++k) A[(24 + i) * 32 + k] = ldexpf(1.f, -(22 + i));
    }
    A[40 * 32 + 0] = 1.f;  A[40 * 32 + 8] = 1.5f * ldexpf(1.f, -23);     // T4 across instructions
    A[41 * 32 + 0] = -1.f; A[41 * 32 + 8] = -1.5f * ldexpf(1.f, -23);
    A[42 * 32 + 0] = 1.f;  A[42 * 32 + 1] = 1.5f * ldexpf(1.f, -23);     // inside one instruction
    A[43 * 32 + 0] = 1.f;  A[43 * 32 + 8] = 1.75f * ldexpf(1.f, -23);
    A[44 * 32 + 0] = 1.f;  A[44 * 32 + 8] = 0.75f * ldexpf(1.f, -23);
    A[45 * 32 + 0] = 1.f;  A[45 * 32 + 1] = 0.75f * ldexpf(1.f, -23);
    A[50 * 32 + 0] = 1.f;                                    // x B row 1 (T1 on the B side)
    // T5: cancellation — large +/- pair and a small addend in one instruction
    A[56 * 32 + 0] = 1.f;  A[56 * 32 + 1] = -1.f; A[56 * 32 + 2] = ldexpf(1.f, -30);
    A[57 * 32 + 0] = 1.f;  A[57 * 32 + 1] = ldexpf(1.f, -30); A[57 * 32 + 8] = -1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 16384 + N * 128 + 2048;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(dA, dB, N, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    auto ulps = [&](int m, int n) { return (double)(fabsf(D[m * N + n]) - 1.f) * 8388608.0; };
    printf("T1 operand low bits: A=1+2^-11+2^-13 -> %.10f (A side), %.10f (negated), %.10f (B side); just below half ulp -> %.10f\n",
           D[0 * N + 0], D[2 * N + 0], D[50 * N + 1], D[3 * N + 0]);
    printf("   (1.0 = truncation, 1.0009765625 = round to nearest)\n");
    for (int i = 0; i < 10; ++i)
        printf("T2 one instruction : 1 + 7 x 2^-%d -> D-1 = %.3f ulp (exact %.4f)\n", 22 + i, ulps(8 + i, 0), 7.0 * ldexp(1.0, 1 - i));
    for (int i = 0; i < 10; ++i)
        printf("T3 next instruction: 1 + 8 x 2^-%d -> D-1 = %.3f ulp (exact %.4f)\n", 22 + i, ulps(24 + i, 0), 8.0 * ldexp(1.0, 1 - i));
    printf("T4 write-back: 1 + 1.5 ulp -> %.3f (next instr), -(1 + 1.5 ulp) -> %.3f, same instr -> %.3f; 1 + 1.75 ulp -> %.3f; "
           "1 + 0.75 ulp -> %.3f (next), %.3f (same)\n", ulps(40, 0), ulps(41, 0), ulps(42, 0), ulps(43, 0), ulps(44, 0), ulps(45, 0));
    printf("T5 cancellation: 1 - 1 + 2^-30 in one instruction -> %.3e (exact 9.313e-10); 1 + 2^-30 then - 1 -> %.3e\n",
           D[56 * N + 0], D[57 * N + 0]);
    return 0;
}
