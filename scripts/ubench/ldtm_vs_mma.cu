// Microbenchmark: TMEM -> register read rate of the K2T epilogue pattern (16 warps, each one 56-column
// tcgen05.ld group = x32 + x16 + x8, then wait) with and without mxf4 MMAs running concurrently into the
// other accumulator buffer.  Answers: is the epilogue's accumulator read (128 lanes x 224 columns x 4 B per
// tile) bound by TMEM read bandwidth, and does it slow down under MMA load?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I np-sims_b200/csrc scripts/ubench/ldtm_vs_mma.cu -o /tmp/ldtm_vs_mma
#include <cstdio>
#include "tc.cuh"
using namespace nps;
constexpr int kWarps = 17;   // 16 readers + 1 MMA issuer
__global__ void __launch_bounds__(kWarps * 32, 1) k(uint32_t iters, uint32_t with_mma, uint32_t cols_per_warp, unsigned long long* out, uint32_t* sink) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    __shared__ volatile uint32_t stop;
    for (uint32_t i = threadIdx.x; i < (16384 + 224 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x2A2A2A2Au;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 4) {
        for (uint32_t c = 448; c < 512u; c += 8u) tmem_st8_fill(tmem + ((warp * 32u) << 16) + c, 0x7F7F7F7Fu);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 16) {
        if (with_mma && lane == 0) {
            const uint32_t idesc = umma_idesc_mxf4(128, 224);
            uint32_t n = 0;
            while (!stop) {
                for (int j = 0; j < 11; ++j)
                    umma_mxf4(tmem + 224, umma_desc_k_sw128(base + (j & 3) * 32), umma_desc_k_sw128(base + 16384 + (j & 3) * 32), idesc, tmem + 448, tmem + 480, j != 0);
                umma_commit(&bar);
                mbar_wait(&bar, n & 1u);
                ++n;
            }
            out[32] = n;
        }
    } else {
        uint32_t acc = 0;
        const uint32_t t_addr = tmem + (((warp & 3) * 32u) << 16) + (warp >> 2) * 56u;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const long long t0 = clock64();
        for (uint32_t it = 0; it < iters; ++it) {
            uint32_t v[56];
            tmem_ld32_at<0>(t_addr, v);
            if (cols_per_warp > 32) tmem_ld16_at<32>(t_addr + 32u, v);
            if (cols_per_warp > 48) tmem_ld8_at<48>(t_addr + 48u, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 56; j += 2) acc |= v[j] & v[j + 1];
        }
        const long long t1 = clock64();
        if (acc == 0x12345) sink[0] = acc;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == 0) stop = 1;
        if (lane == 0) out[warp] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
int main() {
    unsigned long long* d; uint32_t* s; cudaMalloc(&d, 64 * 8); cudaMalloc(&s, 4);
    const size_t smem = 16384 + 224 * 128 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (uint32_t cols : {32u, 56u})
        for (uint32_t with_mma : {0u, 1u}) {
            const uint32_t iters = 2000;
            k<<<1, kWarps * 32, smem>>>(iters, with_mma, cols, d, s);
            k<<<1, kWarps * 32, smem>>>(iters, with_mma, cols, d, s);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            unsigned long long mx = 0; for (int i = 0; i < 16; ++i) mx = h[i] > mx ? h[i] : mx;
            const double bytes = 16.0 * iters * cols * 32 * 4;
            printf("cols/warp=%u mma=%u: %.0f cycles per 16-warp read of 128 x %u x 4 B (%.1f B/clk/SM)%s\n", cols, with_mma, (double)mx / iters, cols * 4, bytes / mx,
                   with_mma ? "" : "");
            if (with_mma) printf("    concurrent MMA batches (11 x M128 N224 K64 mxf4): %llu in %llu cycles = %.0f cycles per batch (1232 alone)\n", h[32], mx, (double)mx / h[32]);
        }
    return 0;
}
