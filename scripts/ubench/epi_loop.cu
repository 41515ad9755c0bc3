// Microbenchmark: the epilogue's fast path in isolation (8 warps: TMEM load one group ahead,
// 16 three-input ANDs, one warp reduce) -- cycles per 32-column group.
#include <cstdio>
#include "tc.cuh"
using namespace nps;
__global__ void __launch_bounds__(512, 1) k_epi(uint32_t tiles, uint32_t groups, int mode, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t tptr;
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr, warp = threadIdx.x >> 5;
    const uint32_t t_addr = tmem + (((warp & 3) * 32u) << 16) + (warp >> 2) * 128u;
    uint32_t found = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (uint32_t t = 0; t < tiles; ++t) {
        uint32_t va[32], vb[32];
        auto test = [&](const uint32_t (&v)[32]) -> uint32_t {
            uint32_t m = 0;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                uint32_t acc = v[8 * qd] & v[8 * qd + 1];
#pragma unroll
                for (int j = 2; j < 8; j += 2) acc &= v[8 * qd + j] & v[8 * qd + j + 1];
                m |= (acc >> 31 ^ 1u) << qd;
            }
            return m;
        };
        uint32_t g = 0;
        tmem_ld32(t_addr, va);
        tmem_ld_wait();
        while (g < groups) {
            if (g + 1 < groups) tmem_ld32(t_addr + ((g + 1) & 3) * 32u, vb);
            const uint32_t ma = test(va);
            if (g + 1 < groups) tmem_ld_wait();
            if (mode == 0) found += __reduce_or_sync(0xffffffffu, ma ^ 15u) & 16u; else found += ma & 16u;
            if (g + 1 >= groups) break;
            if (g + 2 < groups) tmem_ld32(t_addr + ((g + 2) & 3) * 32u, va);
            const uint32_t mb = test(vb);
            if (g + 2 < groups) tmem_ld_wait();
            if (mode == 0) found += __reduce_or_sync(0xffffffffu, mb ^ 15u) & 16u; else found += mb & 16u;
            g += 2;
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
    for (int mode : {0, 1})
        for (int warps : {8, 16}) {
            const uint32_t tiles = 500, groups = 4;
            k_epi<<<1, warps * 32>>>(tiles, groups, mode, d, s);
            k_epi<<<1, warps * 32>>>(tiles, groups, mode, d, s);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            unsigned long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("mode=%d warps=%2d: %.1f cycles per group per warp, %.1f cycles per tile (4 groups)\n", mode, warps, (double)mx / (tiles * groups), (double)mx / tiles);
        }
    return 0;
}
