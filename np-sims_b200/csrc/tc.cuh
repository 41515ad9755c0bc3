// tc.cuh — tcgen05 / TMEM / UMMA-descriptor helpers (sm_100a inline PTX) shared by the
// tensor-core kernels (mma_scan.cu: batched Hamming scan; project_tc.cu: sign(X.W^T) hashing).
#pragma once
#include "common.cuh"

namespace nps {

// ---- TMEM allocation (one full warp executes these) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- shared-memory matrix descriptor: K-major operand, 128-byte swizzle ----------------------
// Canonical layout (cute::UMMA K-major SW128): rows of 128 bytes, 8-row groups of 1024 bytes
// (stride byte offset), 16-byte chunk c of row r stored at chunk position c ^ (r & 7).
// The tile base must be 1024-byte aligned; advancing along K inside the 128-byte row is done
// by adding the byte offset to the start address.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for SW128 K-major) = 16 B
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset = 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
    return d;
}

// K-major operand WITHOUT swizzle (cute::UMMA INTERLEAVE): 8-row x 16-byte core matrices stored
// contiguously (128 B); lbo = byte distance between the two 16-byte K chunks of one MMA, sbo =
// byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_k_noswizzle(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// ---- instruction descriptor (kind::f8f6f4 / kind::tf32 / kind::f16), FP32 accumulate ----------
// a_fmt/b_fmt: f8f6f4: 0 = E4M3, 1 = E5M2;  tf32 kind: 2 = TF32;  f16 kind: 0 = F16, 1 = BF16
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t M, uint32_t N) {
    return (1u << 4)              // D format: F32
           | (a_fmt << 7) | (b_fmt << 10)
           | (0u << 15) | (0u << 16)   // A, B K-major
           | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::mxf4: E2M1 (4-bit) operands packed two per byte, K = 64 per instruction, block-32 UE8M0
// scale factors read from TMEM.  Instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled):
__host__ __device__ constexpr uint32_t umma_idesc_mxf4(uint32_t M, uint32_t N) {
    return (1u << 7) | (1u << 10)            // A, B format: E2M1
           | ((N >> 3) << 17) | (1u << 23)   // N; scale format UE8M0
           | ((M >> 4) << 24);               // M; scale-factor ids 0; K = 64
}
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 8 consecutive columns, all set to v
__device__ __forceinline__ void tmem_st8_fill(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// all MMAs issued so far by this thread -> arrive on an mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns -------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// this warp's 32 lanes x 64 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                 : "r"(taddr)
                 : "memory");
}
// this warp's 32 lanes x 32 consecutive 32-bit columns into v[O .. O+32)
template <int O, int LEN>
__device__ __forceinline__ void tmem_ld32_at(uint32_t taddr, uint32_t (&v)[LEN]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[O + 0]), "=r"(v[O + 1]), "=r"(v[O + 2]), "=r"(v[O + 3]), "=r"(v[O + 4]), "=r"(v[O + 5]), "=r"(v[O + 6]), "=r"(v[O + 7]), "=r"(v[O + 8]), "=r"(v[O + 9]), "=r"(v[O + 10]), "=r"(v[O + 11]), "=r"(v[O + 12]), "=r"(v[O + 13]), "=r"(v[O + 14]), "=r"(v[O + 15]), "=r"(v[O + 16]), "=r"(v[O + 17]), "=r"(v[O + 18]), "=r"(v[O + 19]), "=r"(v[O + 20]), "=r"(v[O + 21]), "=r"(v[O + 22]), "=r"(v[O + 23]), "=r"(v[O + 24]), "=r"(v[O + 25]), "=r"(v[O + 26]), "=r"(v[O + 27]), "=r"(v[O + 28]), "=r"(v[O + 29]), "=r"(v[O + 30]), "=r"(v[O + 31])
                 : "r"(taddr)
                 : "memory");
}
// this warp's 32 lanes x 16 consecutive 32-bit columns into v[O .. O+16)
template <int O, int LEN>
__device__ __forceinline__ void tmem_ld16_at(uint32_t taddr, uint32_t (&v)[LEN]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[O + 0]), "=r"(v[O + 1]), "=r"(v[O + 2]), "=r"(v[O + 3]), "=r"(v[O + 4]), "=r"(v[O + 5]), "=r"(v[O + 6]), "=r"(v[O + 7]), "=r"(v[O + 8]), "=r"(v[O + 9]), "=r"(v[O + 10]), "=r"(v[O + 11]), "=r"(v[O + 12]), "=r"(v[O + 13]), "=r"(v[O + 14]), "=r"(v[O + 15])
                 : "r"(taddr)
                 : "memory");
}
// this warp's 32 lanes x 8 consecutive 32-bit columns into v[O .. O+8)
template <int O, int LEN>
__device__ __forceinline__ void tmem_ld8_at(uint32_t taddr, uint32_t (&v)[LEN]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[O + 0]), "=r"(v[O + 1]), "=r"(v[O + 2]), "=r"(v[O + 3]), "=r"(v[O + 4]), "=r"(v[O + 5]), "=r"(v[O + 6]), "=r"(v[O + 7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Programmatic dependent launch: a kernel launched with programmaticStreamSerializationAllowed may
// start (prologue: barrier init, TMEM allocation) while its predecessor in the stream is still
// draining; pdl_wait() blocks until the predecessor grid has completed and its writes are visible
// (a no-op without the launch attribute).  pdl_launch_dependents() lets the successor start early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of one TPC issue ONE MMA of
// M = 256; each CTA supplies its 128 A rows and HALF of the B rows from its own shared memory and
// keeps its 128 accumulator rows in its own TMEM.  Only the leader (cluster rank 0) issues.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in the LEADER CTA (rank 0) of the pair.
// Default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id): a
// .release.cluster arrive is a cluster-wide fence that drains every outstanding store of the thread —
// measured ~1300 cycles per arrive here, against ~90 for this form.  What the leader's MMA reads from
// this CTA's shared memory has been handed to the async proxy by fence.proxy.async before the arrive;
// TMEM reads have completed (tcgen05.wait::ld) before an accumulator is released.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// wait on a local mbarrier whose arrivals may come from the other CTA of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void umma_mxf4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// all MMAs issued so far by this thread -> arrive on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

}  // namespace nps
