// Microbenchmark: per-instruction issue timestamps of tcgen05.mma / tcgen05.commit from one thread.
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"
using namespace nps;

__global__ void __launch_bounds__(128, 1) k_trace(uint32_t N, uint32_t groups, uint32_t per_commit, long long* ts) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    __shared__ uint64_t bar[64];
    __shared__ uint32_t tptr;
    for (uint32_t i = threadIdx.x; i < (16384 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(0, 0, 128, N);
        const uint64_t da = umma_desc_k_sw128(base), db = umma_desc_k_sw128(base + 16384);
        umma_f8(tmem, da, db, idesc, 0);
        umma_commit(&bar[63]); mbar_wait(&bar[63], 0);
        int n = 0;
        ts[n++] = clock64();
        for (uint32_t g = 0; g < groups; ++g) {
            for (uint32_t j = 0; j < per_commit; ++j) { umma_f8(tmem + (g & 1) * N, da + 2 * (j & 3), db + 2 * (j & 3), idesc, j != 0); ts[n++] = clock64(); }
            umma_commit(&bar[g]); ts[n++] = clock64();
        }
        // completion time of every group
        for (uint32_t g = 0; g < groups; ++g) { mbar_wait(&bar[g], 0); ts[n++] = clock64(); }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
    long long* d; cudaMalloc(&d, 4096 * 8);
    const size_t smem = 16384 + 256 * 128 + 2048;
    cudaFuncSetAttribute(k_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (uint32_t N : {256u, 128u})
        for (uint32_t pc : {4u, 8u}) {
            const uint32_t groups = 12;
            k_trace<<<1, 128, smem>>>(N, groups, pc, d);
            k_trace<<<1, 128, smem>>>(N, groups, pc, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            static long long h[4096]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("N=%u per_commit=%u  issue times (cycles since start): ", N, pc);
            int n = 1;
            for (uint32_t g = 0; g < groups; ++g) { printf("| "); for (uint32_t j = 0; j <= pc; ++j) printf("%lld ", h[n++] - h[0]); }
            printf("\n   group completion observed at: ");
            for (uint32_t g = 0; g < groups; ++g) printf("%lld ", h[n++] - h[0]);
            printf("\n"); fflush(stdout);
        }
    return 0;
}
