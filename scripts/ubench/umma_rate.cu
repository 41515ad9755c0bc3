// Microbenchmark (development + roofline denominator): issue rate of tcgen05.mma kind::f8f6f4,
// M=128, SS operands (K-major SWIZZLE_128B), one CTA per SM, back-to-back from one thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I np-sims_b200/csrc scripts/ubench/umma_rate.cu -o umma_rate
#include <cstdio>
#include <cstdlib>
#include "tc.cuh"
using namespace nps;

__global__ void __launch_bounds__(128, 1) k_rate(uint32_t N, uint32_t iters, uint32_t per_commit, uint32_t mode, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t tptr;
    for (uint32_t i = threadIdx.x; i < (16384 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(0, 0, 128, N);
        const uint32_t a = base, b = base + 16384;
        uint32_t ph = 0;
        // warm-up
        umma_f8(tmem, umma_desc_k_sw128(a), umma_desc_k_sw128(b), idesc, 0);
        umma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1;
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; it += per_commit) {
            for (uint32_t j = 0; j < per_commit; ++j)
                umma_f8(tmem + ((it / per_commit) & 1) * N, umma_desc_k_sw128(a + (j & 3) * 32), umma_desc_k_sw128(b + (j & 3) * 32), idesc, j != 0);
            if (it + per_commit < iters) {
                umma_commit(&bar2);   // intermediate commits: nobody waits on these
                if (mode & 1u) umma_commit(&bar3);                       // a second commit right behind
                if (mode & 2u) { mbar_wait(&bar3, 1u); tc_fence_after(); }   // an already-complete wait, like a_full
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, ph);
        const long long t1 = clock64();
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
    unsigned long long* d; cudaMalloc(&d, 148 * 8);
    const size_t smem = 16384 + 256 * 128 + 2048;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {148})
        for (uint32_t N : {128u, 256u})
            for (uint32_t cfg : {0x0004u, 0x0008u, 0x0014u, 0x0800u, 0x10004u, 0x20004u, 0x20008u}) {
                const uint32_t pc = cfg & 0xFFFFu, mode = cfg >> 16;
                const uint32_t iters = 2040 / pc * pc;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                k_rate<<<grid, 128, smem>>>(N, iters, pc, mode, d);
                cudaEventRecord(e0);
                k_rate<<<grid, 128, smem>>>(N, iters, pc, mode, d);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                unsigned long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i]; cyc /= grid;
                const double mac = 128.0 * N * 32 * iters;
                printf("grid=%3d N=%3u mode=%u mmas/commit=%4u: %.1f cycles/MMA, %.0f MAC/clk/SM, kernel %.3f ms -> %.2f PFLOP/s chip\n",
                       grid, N, mode, pc, cyc / iters, mac / cyc, ms, 2.0 * mac * grid / (ms * 1e-3) / 1e15);
                fflush(stdout);
            }
    return 0;
}
