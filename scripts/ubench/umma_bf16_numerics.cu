// Probe: accumulation numerics of tcgen05.mma kind::f16 with BF16 operands (K = 16 per instruction) —
// the same questions as umma_tf32_numerics.cu T2-T5, for the bf16 x 2 variant of project_tc.cu.
// D[m][0] = sum_k A[m][k] * 1; A rows hand-made from exactly representable bf16 values (powers of two).
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_bf16.h>
#include "tc.cuh"
using namespace nps;

__global__ void __launch_bounds__(128, 1) k_probe(const __nv_bfloat16* A, const __nv_bfloat16* B, uint32_t N, float* D) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    auto put = [&](uint8_t* tile, const __nv_bfloat16* src, uint32_t rows) {   // row-major [rows][64 bf16] -> SWIZZLE_128B
        for (uint32_t i = threadIdx.x; i < rows * 8; i += blockDim.x) {
            const uint32_t r = i >> 3, c = i & 7;
            *reinterpret_cast<uint4*>(tile + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) =
                *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(src) + r * 128 + c * 16);
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
        const uint32_t idesc = umma_idesc(1, 1, 128, N);   // BF16 x BF16 -> F32
        const uint64_t da = umma_desc_k_sw128(base), db = umma_desc_k_sw128(base + 16384);
        for (int j = 0; j < 4; ++j) umma_f16(tmem, da + 2 * j, db + 2 * j, idesc, j != 0);   // K = 16 each
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
    std::vector<float> Af(128 * 64, 0.f), Bf(N * 64, 0.f);
    for (int k = 0; k < 64; ++k) Bf[k] = 1.f;
    for (int i = 0; i < 10; ++i) {                           // T2: 1 + 15 * 2^-(22+i) inside one instruction (K = 16)
        Af[(8 + i) * 64 + 0] = 1.f;
        for (int k = 1; k < 16; ++k) Af[(8 + i) * 64 + k] = ldexpf(1.f, -(22 + i));
    }
    for (int i = 0; i < 10; ++i) {                           // T3: 1 (instr 0) + 16 * 2^-(22+i) (instr 1)
        Af[(24 + i) * 64 + 0] = 1.f;
        for (int k = 16; k < 32; ++k) Af[(24 + i) * 64 + k] = ldexpf(1.f, -(22 + i));
    }
    Af[40 * 64 + 0] = 1.f;  Af[40 * 64 + 16] = 1.5f * ldexpf(1.f, -23);     // T4 write-back rounding
    Af[41 * 64 + 0] = -1.f; Af[41 * 64 + 16] = -1.5f * ldexpf(1.f, -23);
    Af[42 * 64 + 0] = 1.f;  Af[42 * 64 + 1] = 1.5f * ldexpf(1.f, -23);
    Af[56 * 64 + 0] = 1.f;  Af[56 * 64 + 1] = -1.f; Af[56 * 64 + 2] = ldexpf(1.f, -30);   // T5 cancellation
    std::vector<__nv_bfloat16> A(Af.size()), B(Bf.size());
    for (size_t i = 0; i < Af.size(); ++i) A[i] = __float2bfloat16(Af[i]);
    for (size_t i = 0; i < Bf.size(); ++i) B[i] = __float2bfloat16(Bf[i]);
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 16384 + N * 128 + 2048;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(dA, dB, N, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    auto ulps = [&](int m) { return (double)(fabsf(D[m * N]) - 1.f) * 8388608.0; };
    for (int i = 0; i < 10; ++i)
        printf("T2 one instruction : 1 + 15 x 2^-%d -> D-1 = %.3f ulp (exact %.4f)\n", 22 + i, ulps(8 + i), 15.0 * ldexp(1.0, 1 - i));
    for (int i = 0; i < 10; ++i)
        printf("T3 next instruction: 1 + 16 x 2^-%d -> D-1 = %.3f ulp (exact %.4f)\n", 22 + i, ulps(24 + i), 16.0 * ldexp(1.0, 1 - i));
    printf("T4 write-back: 1 + 1.5 ulp -> %.3f (next instr), -(1 + 1.5 ulp) -> %.3f, same instr -> %.3f\n", ulps(40), ulps(41), ulps(42));
    printf("T5 cancellation: 1 - 1 + 2^-30 in one instruction -> %.3e (exact 9.313e-10)\n", D[56 * N]);
    return 0;
}
