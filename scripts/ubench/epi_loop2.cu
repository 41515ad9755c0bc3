// Microbenchmark: epilogue fast-path variants -- one tcgen05.ld.x64 per warp per tile (W warps).
#include <cstdio>
#include "tc.cuh"
using namespace nps;
__global__ void __launch_bounds__(512, 1) k_epi(uint32_t tiles, uint32_t loads, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t tptr;
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr, warp = threadIdx.x >> 5;
    const uint32_t t_addr = tmem + (((warp & 3) * 32u) << 16) + ((warp >> 2) * 64u & 255u);
    uint32_t found = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t t = 0; t < tiles; ++t) {
        for (uint32_t l = 0; l < loads; ++l) {
            uint32_t v[64];
            tmem_ld64(t_addr + l * 64u, v);
            tmem_ld_wait();
            uint32_t m = 0;
#pragma unroll
            for (int qd = 0; qd < 8; ++qd) {
                uint32_t acc = v[8 * qd] & v[8 * qd + 1];
#pragma unroll
                for (int j = 2; j < 8; j += 2) acc &= v[8 * qd + j] & v[8 * qd + j + 1];
                m |= (acc >> 31 ^ 1u) << qd;
            }
            found += __reduce_or_sync(0xffffffffu, m ^ 255u) & 256u;
        }
        tc_fence_before();
        __syncwarp();
    }
    const long long t1 = clock64();
    if (found == 0x12345) sink[0] = found;
    if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
    unsigned long long* d; uint32_t* s; cudaMalloc(&d, 64 * 8); cudaMalloc(&s, 4);
    for (int warps : {8, 16})
        for (uint32_t loads : {1u, 2u}) {
            const uint32_t tiles = 500;
            k_epi<<<1, warps * 32>>>(tiles, loads, d, s);
            k_epi<<<1, warps * 32>>>(tiles, loads, d, s);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            unsigned long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("warps=%2d x64-loads/tile/warp=%u: %.1f cycles per tile; columns covered per tile = %u\n", warps, loads, (double)mx / tiles, warps / 4 * loads * 64);
        }
    return 0;
}
