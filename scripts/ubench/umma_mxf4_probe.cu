// Probe: tcgen05.mma kind::mxf4 (E2M1 x E2M1, UE8M0 block-32 scales all = 1.0) with +-1 operands.
// Checks D = A.B^T against a host computation for random signs, and times the issue rate.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc.cuh"
using namespace nps;

__device__ __forceinline__ uint32_t idesc_mxf4(uint32_t M, uint32_t N) {
    return (1u << 7) | (1u << 10)        // A, B format: E2M1
           | ((N >> 3) << 17) | (1u << 23)   // scale format UE8M0
           | ((M >> 4) << 24);               // sf ids 0, K = 64
}
__device__ __forceinline__ void umma_mxf4(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}

// A: [128 rows][128 B] swizzled, B: [N rows][128 B] swizzled; raw row-major images are given and swizzled here
__global__ void __launch_bounds__(128, 1) k_probe(const uint8_t* A, const uint8_t* B, uint32_t N, float* D, uint32_t iters, unsigned long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    auto put = [&](uint8_t* tile, const uint8_t* src, uint32_t rows) {
        for (uint32_t i = threadIdx.x; i < rows * 8; i += blockDim.x) {
            const uint32_t r = i >> 3, c = i & 7;
            *reinterpret_cast<uint4*>(tile + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) =
                *reinterpret_cast<const uint4*>(src + r * 128 + c * 16);
        }
    };
    put(smem, A, 128);
    put(smem + 16384, B, N);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr;
    // scale factors: columns [448, 512) of every lane = 0x7F7F7F7F (UE8M0 1.0 in every byte)
    const uint32_t warp = threadIdx.x >> 5;
    for (int c = 0; c < 64; c += 8) tmem_st8(tmem + ((warp * 32u) << 16) + 448 + c, 0x7F7F7F7Fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_mxf4(128, N);
        const uint64_t da = umma_desc_k_sw128(base), db = umma_desc_k_sw128(base + 16384);
        for (int j = 0; j < 4; ++j) umma_mxf4(tmem, da + 2 * j, db + 2 * j, idesc, tmem + 448, tmem + 480, j != 0);
        umma_commit(&bar); mbar_wait(&bar, 0);
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) umma_mxf4(tmem + 256, da + 2 * (it & 3), db + 2 * (it & 3), idesc, tmem + 448, tmem + 480, 1);
        umma_commit(&bar); mbar_wait(&bar, 1);
        cyc[0] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
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
    const uint32_t N = 224;
    std::vector<uint8_t> A(128 * 128), B(N * 128);
    srand(1);
    auto fill = [](std::vector<uint8_t>& v) { for (auto& b : v) { uint8_t lo = (rand() & 1) ? 0xA : 0x2, hi = (rand() & 1) ? 0xA : 0x2; b = lo | (hi << 4); } };
    fill(A); fill(B);
    uint8_t *dA, *dB; float* dD; unsigned long long* dc;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const size_t smem = 16384 + 256 * 128 + 2048;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t iters = 2048;
    k_probe<<<1, 128, smem>>>(dA, dB, N, dD, iters, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N); unsigned long long cyc;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    auto val = [](uint8_t nib) { return (nib & 8) ? -1 : 1; };
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (uint32_t n = 0; n < N; ++n) {
        int dot = 0;
        for (int b = 0; b < 128; ++b) dot += val(A[m * 128 + b] & 15) * val(B[n * 128 + b] & 15) + val(A[m * 128 + b] >> 4) * val(B[n * 128 + b] >> 4);
        if ((float)dot != D[m * N + n]) { if (bad < 5) printf("mismatch m=%d n=%u got %f want %d\n", m, n, D[m * N + n], dot); ++bad; }
    }
    printf("mxf4 probe: N=%u mismatches=%d of %u; D[0]=%f D[1]=%f; %.1f cycles/MMA (M128 x N%u x K64) = %.0f MAC/clk/SM\n", N, bad, 128 * N, D[0], D[1],
           (double)cyc / iters, N, 128.0 * N * 64 * iters / cyc);
    return 0;
}
