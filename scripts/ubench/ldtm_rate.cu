// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld.32x32b.x32 with 4 / 8 / 16 warps.
#include <cstdio>
#include "tc.cuh"
using namespace nps;
__global__ void __launch_bounds__(512, 1) k_ldtm(uint32_t iters, uint32_t cols, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t tptr;
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
        for (uint32_t c = 0; c < cols; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + (((warp & 3) * 32u) << 16) + ((c + (warp >> 2) * 32) & 511u), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc |= v[j];
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345) sink[0] = acc;
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
    unsigned long long* d; uint32_t* s; cudaMalloc(&d, 64 * 8); cudaMalloc(&s, 4);
    for (int warps : {4, 8, 16}) {
        const uint32_t iters = 200, cols = 256;
        k_ldtm<<<1, warps * 32>>>(iters, cols, d, s);
        k_ldtm<<<1, warps * 32>>>(iters, cols, d, s);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        unsigned long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
        const double bytes = (double)warps * iters * (cols / 32) * 32 * 32 * 4;
        printf("warps=%2d: %llu cycles, %.1f cycles per LDTM.x32 per warp, %.1f B/clk/SM\n", warps, mx, (double)mx / (iters * cols / 32), bytes / mx);
    }
    return 0;
}
