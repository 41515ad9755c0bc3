// ref_scan.cuh — K2R: the reference's history-dependent queue (np_sims/ufunc.c:108-195) for ONE
// query at streaming speed.  Backs `hamming_top_10` / `hamming_top_cand` / `lsh.query`
// (ufunc.c:444-551, lsh.py:70-74) bit for bit: ids in queue-SLOT order, tie survivors as the
// sequential insert / evict history leaves them.
//
// The queue is sequential by definition, but almost no row ever enters it: row i is inserted only
// if its distance is strictly below the queue's worst at that moment (ufunc.c:152), and the worst
// never increases.  So any UPPER bound on the worst that is valid for row i gives a superset filter
// that can run in parallel; only the survivors are replayed in row order:
//
//   R0  ref_prefix_kernel  (one CTA per query): distances of the first P rows (P ~ sqrt(N k)), written
//       to the workspace, and tau_P = their k-th smallest distance (shared-memory histogram — a
//       distance is < 64 * words + 1 <= 4096).  For every row >= P the queue has seen those P rows,
//       so its worst is <= tau_P.  Small tables (N <= 32768) are all prefix: this kernel then also
//       replays and is the only launch.
//   R1  ref_scan_kernel    (one persistent CTA per SM, same TMA ring + register-tile XOR+POPC sweep
//       as K2, scan.cuh): streams rows [P, N) once at HBM speed and keeps (row, distance) of rows with
//       distance < tau_P — an expected (N - P) k / P ~ sqrt(N k) of them.  Each row chunk stages its
//       survivors in shared memory, sorts them by row and writes one ordered segment.
//   R2  replay_ordered     (device function; run by the LAST CTA of R1 to finish — no third launch):
//       the reference queue itself, in shared memory, over the P prefix distances and then the
//       segments in chunk order.  One warp tests 32 candidates at a time against the live worst and
//       inserts the few that pass one by one (overwrite the FIRST slot holding the maximum,
//       ufunc.c:163-169).
//
// If a chunk's segment overflows (rows laid out adversarially, e.g. sorted by similarity to the
// query so that the prefix is not representative) a flag is raised and the caller falls back to
// K5 + the dense replay (replay.cu) — exact either way.
#pragma once
#include "common.cuh"

namespace nps {

constexpr uint32_t kRefMaxWords = 63;        // distances < 4096: the histogram of R0
constexpr uint32_t kRefBins = 4096;
constexpr uint32_t kRefSmallRows = 32768;    // tables up to here are all prefix (one launch)
constexpr uint32_t kRefMaxPrefix = 32768;
constexpr uint32_t kRefMaxCap = 4096;        // entries per chunk segment
constexpr uint32_t kRefMaxChunks = 4096;
constexpr uint32_t kRefBatch = 4096;         // replay batch (entries / 4 x distances) staged in shared memory
constexpr int kRefPrefixThreads = 512;        // R0: one prefix row per thread (512: 128 registers for the replay)

struct RefParams {
    const uint64_t* hashes;    // [N, W]
    const uint64_t* queries;   // [nq, W]
    uint64_t* out_rows;        // [nq, k] slot order, NPS_NONE where never filled
    uint64_t* out_scores;      // [nq, k] nullable
    uint32_t* any_overflow;    // [1] set when any query's segment overflowed (gate of the fallback launches)
    uint32_t* ctrl;            // [nq, 4] control words (kRefCtrl*)
    uint32_t* hist;            // [nq, kRefBins] distance histogram of the prefix rows (zero between calls)
    uint8_t* ws;               // per-query blocks of ws_stride bytes
    unsigned long long ws_stride;
    unsigned long long N;
    uint32_t words, k;
    uint32_t P;                // prefix rows (R0); P == N: no R1
    uint32_t tile_rows;        // R1 tile (generic-width kernel; the specialised one knows it)
    uint32_t chunk_tiles, num_chunks, cap, stages;
    uint32_t off_pre, off_cnt, off_seg;   // byte offsets inside a query's block
};
// control words of a query.  The done counters and the histogram are left at zero by the call itself
// (the last CTA cleans up), so a workspace only has to be zeroed once (ref_zero_bytes()).
enum { kRefCtrlTau = 0, kRefCtrlDone = 1, kRefCtrlOverflow = 2, kRefCtrlPrefixDone = 3 };

// shared memory the replay needs (slots + one batch)
__host__ __device__ inline uint32_t ref_replay_smem(uint32_t k, uint32_t num_chunks) {
    return ((k * 8 + 15u) & ~15u) + kRefBatch * 8 + (((num_chunks + 1) * 4 + 15u) & ~15u) + k * 4 + 64;
}

// ---------------------------------------------------------------------------------------------
// The reference queue, held in the REGISTERS of warp 0 of the calling thread group.
// Slot s is ONE 32-bit key  score << 11 | (2047 - s)  (score < 4096, s < 2048; a never-filled slot
// has the all-ones score), so "the FIRST slot holding the maximum" (ufunc.c:163-169: strict '>'
// keeps the lowest slot) is a plain integer maximum.  Lane l owns slots l, l + 32, ... (SPL of them,
// SPL = 1 for the reference's k = 10) in registers and caches their maximum; one redux.sync gives the
// queue's worst and its slot.  An insert = one shuffle of the candidate's distance, a predicated
// register update + re-scan of SPL registers by the owner lane, the redux — no shared-memory round
// trip on the serial path.  Row ids of the slots live in shared memory (written on insert, read once
// at the end).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kRefSlotBits = 11;
constexpr uint32_t kRefSlotMask = (1u << kRefSlotBits) - 1u;
constexpr uint32_t kRefEmptyScore = 0xFFFFFFFFu >> kRefSlotBits;

template <int SPL>
struct ReplayQueue {
    uint32_t key[SPL];          // key[j]: slot j * 32 + lane (0 for slots >= k)
    uint64_t* row;              // [k] shared
    uint32_t lane_best;         // max over this lane's slots
    uint32_t worst, worst_slot; // warp-uniform; worst == kRefEmptyScore: the queue is not full

    __device__ __forceinline__ void rescan() {
        uint32_t best = 0;
#pragma unroll
        for (int j = 0; j < SPL; ++j) best = key[j] > best ? key[j] : best;
        lane_best = best;
    }
    __device__ __forceinline__ void publish() {
        const uint32_t best = __reduce_max_sync(0xffffffffu, lane_best);
        worst = best >> kRefSlotBits;
        worst_slot = kRefSlotMask - (best & kRefSlotMask);
    }
    // up to 32 candidates in row order, one per lane: each is tested against the LIVE worst
    // (strict '<', ufunc.c:152); d < 4096
    __device__ __forceinline__ void group(uint32_t d, uint64_t r, bool valid, int lane) {
        uint32_t pend = __ballot_sync(0xffffffffu, valid && d < worst);
        while (pend) {
            const int src = __ffs(pend) - 1;
            const uint32_t slot = worst_slot;
            const uint32_t nk = (__shfl_sync(0xffffffffu, d, src) << kRefSlotBits) | (kRefSlotMask - slot);
            if (lane == src) row[slot] = r;
            const bool owner = (slot & 31u) == (uint32_t)lane;
            const int js = (int)(slot >> 5);
#pragma unroll
            for (int j = 0; j < SPL; ++j) key[j] = (owner && j == js) ? nk : key[j];
            rescan();
            publish();
            pend &= ~((2u << src) - 1u);
            pend &= __ballot_sync(0xffffffffu, valid && d < worst);
        }
    }
};

struct ReplayArgs {
    uint8_t* blk;               // the query's workspace block
    uint64_t* out_rows;         // [k]
    uint64_t* out_scores;       // [k] nullable
    uint32_t k, P, num_chunks, cap, off_pre, off_cnt, off_seg;
};

// Replay by `nthreads` threads (a multiple of 32, warp 0 included) that meet on named barrier `bar`.
// smem: ref_replay_smem(k, num_chunks) bytes, 16-byte aligned.
template <int SPL>
__device__ __forceinline__ void replay_ordered(const ReplayArgs& a, uint8_t* smem, int tid, int nthreads, uint32_t bar) {
    const uint32_t k = a.k;
    uint64_t* row_s = reinterpret_cast<uint64_t*>(smem);                        // [k]
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem + ((k * 8 + 15u) & ~15u)); // [kRefBatch]
    uint32_t* offs = reinterpret_cast<uint32_t*>(buf + kRefBatch);              // [num_chunks + 1]
    uint32_t* key_out = offs + ((a.num_chunks + 1 + 3u) & ~3u);                 // [k] final keys
    const uint16_t* pre = reinterpret_cast<const uint16_t*>(a.blk + a.off_pre); // 16-byte aligned
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(a.blk + a.off_cnt);
    const uint64_t* seg = reinterpret_cast<const uint64_t*>(a.blk + a.off_seg);
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t P = a.P, NC = a.num_chunks;

    // ---- fill (ufunc.c:153-157): the first min(k, N) rows take slots 0.. in order
    const uint32_t filled = P < k ? P : k;
    for (uint32_t i = tid; i < k; i += nthreads) row_s[i] = i < filled ? (uint64_t)i : kNone;
    ReplayQueue<SPL> Q;
    Q.row = row_s;
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const uint32_t s = (uint32_t)j * 32u + lane;
            const uint32_t sc = s < filled ? (uint32_t)__ldcg(pre + s) : kRefEmptyScore;
            Q.key[j] = s < k ? ((sc << kRefSlotBits) | (kRefSlotMask - s)) : 0u;
        }
        Q.rescan();
        Q.publish();
    }
    // segment offsets: counts -> shared (all threads, independent loads), exclusive scan by warp 0
    for (uint32_t c0 = 0; c0 < NC; c0 += 4u * nthreads) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = c0 + j * nthreads + tid;
            if (c < NC) v[j] = __ldcg(cnt + c);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = c0 + j * nthreads + tid;
            if (c < NC) offs[c + 1] = v[j];
        }
    }
    named_bar_sync_rt(bar, nthreads);
    if (warp == 0) {
        uint32_t run = 0;
        for (uint32_t c0 = 0; c0 < NC; c0 += 32) {
            const uint32_t c = c0 + lane;
            const uint32_t v = c < NC ? offs[c + 1] : 0u;
            uint32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            __syncwarp();
            if (c < NC) offs[c + 1] = run + inc;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) offs[0] = 0;
    }
    named_bar_sync_rt(bar, nthreads);

    // ---- phase A: prefix rows [filled, P), dense distances, 4 * kRefBatch per batch (16-byte loads)
    uint16_t* dbuf = reinterpret_cast<uint16_t*>(buf);
    constexpr uint32_t kABatch = kRefBatch * 4;
    for (uint32_t base = filled & ~7u; base < P; base += kABatch) {   // base stays a multiple of 8
        const uint32_t n = (P - base) < kABatch ? (P - base) : kABatch;
        const uint32_t nv = (n + 7u) >> 3;                            // reads up to 14 bytes past P: inside the block
        const uint4* src = reinterpret_cast<const uint4*>(pre + base);
        uint4* dst = reinterpret_cast<uint4*>(dbuf);
        for (uint32_t i0 = 0; i0 < nv; i0 += 4u * nthreads) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t i = i0 + j * nthreads + tid;
                if (i < nv) v[j] = __ldcg(src + i);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t i = i0 + j * nthreads + tid;
                if (i < nv) dst[i] = v[j];
            }
        }
        named_bar_sync_rt(bar, nthreads);
        if (warp == 0) {
            for (uint32_t i = 0; i < n; i += 32) {
                const uint32_t e = i + lane;
                const bool valid = e < n && base + e >= filled;
                const uint32_t d = e < n ? (uint32_t)dbuf[e] : 0u;
                Q.group(d, (uint64_t)(base + e), valid, lane);
            }
        }
        named_bar_sync_rt(bar, nthreads);
    }

    // ---- phase B: the ordered segments of rows [P, N)
    const uint32_t M = offs[NC];
    for (uint32_t base = 0; base < M; base += kRefBatch) {
        const uint32_t n = (M - base) < kRefBatch ? (M - base) : kRefBatch;
        for (uint32_t i0 = 0; i0 < n; i0 += 4u * nthreads) {
            uint64_t v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t i = i0 + j * nthreads + tid;
                if (i < n) {
                    const uint32_t e = base + i;
                    uint32_t lo = 0, hi = NC;            // largest c with offs[c] <= e
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (offs[mid] <= e) lo = mid;
                        else hi = mid;
                    }
                    v[j] = __ldcg(seg + (size_t)lo * a.cap + (e - offs[lo]));
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t i = i0 + j * nthreads + tid;
                if (i < n) buf[i] = v[j];
            }
        }
        named_bar_sync_rt(bar, nthreads);
        if (warp == 0) {
            for (uint32_t i = 0; i < n; i += 32) {
                const uint32_t e = i + lane;
                const uint64_t v = e < n ? buf[e] : 0ull;
                Q.group((uint32_t)(v & 0xFFFFu), v >> 16, e < n, lane);
            }
        }
        named_bar_sync_rt(bar, nthreads);
    }
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const uint32_t s = (uint32_t)j * 32u + lane;
            if (s < k) key_out[s] = Q.key[j];
        }
    }
    named_bar_sync_rt(bar, nthreads);
    for (uint32_t i = tid; i < k; i += nthreads) {
        const uint64_t r = row_s[i];
        a.out_rows[i] = r;
        if (a.out_scores) a.out_scores[i] = r == kNone ? kNone : (uint64_t)(key_out[i] >> kRefSlotBits);
    }
}

#ifdef NPS_REF_DEVICE_CODE
// One out-of-line copy per translation unit (not one per width-specialised kernel): cold code, run once.
__device__ __noinline__ static void replay_run(ReplayArgs a, uint8_t* smem, int tid, int nthreads, uint32_t bar) {
    if (a.k <= 32) replay_ordered<1>(a, smem, tid, nthreads, bar);
    else if (a.k <= 128) replay_ordered<4>(a, smem, tid, nthreads, bar);
    else if (a.k <= 512) replay_ordered<16>(a, smem, tid, nthreads, bar);
    else if (a.k <= 1024) replay_ordered<32>(a, smem, tid, nthreads, bar);
    else replay_ordered<64>(a, smem, tid, nthreads, bar);
}
__device__ __forceinline__ ReplayArgs replay_args(const RefParams& p, uint32_t qi) {
    ReplayArgs a;
    a.blk = p.ws + (size_t)qi * p.ws_stride;
    a.out_rows = p.out_rows + (size_t)qi * p.k;
    a.out_scores = p.out_scores ? p.out_scores + (size_t)qi * p.k : nullptr;
    a.k = p.k;
    a.P = p.P;
    a.num_chunks = p.num_chunks;
    a.cap = p.cap;
    a.off_pre = p.off_pre;
    a.off_cnt = p.off_cnt;
    a.off_seg = p.off_seg;
    return a;
}

// ---------------------------------------------------------------------------------------------
// R1 shared-memory layout
// ---------------------------------------------------------------------------------------------
struct RefSmem {
    uint32_t stage_off, q_off, emit_off, misc_off, bar_off, total;
};
__host__ __device__ inline uint32_t ref_q_stride(uint32_t words) { return ((words + 1) / 2 + 8) * 16; }
__host__ __device__ inline RefSmem ref_smem_layout(uint32_t tile_bytes, uint32_t stages, uint32_t words, uint32_t cap,
                                                   uint32_t k, uint32_t num_chunks) {
    RefSmem s;
    uint32_t off = 0;
    s.stage_off = off;
    off += stages * ((tile_bytes + 127u) & ~127u);
    s.q_off = off;
    off += ref_q_stride(words);
    s.emit_off = off;
    off += cap * 8;
    s.misc_off = off;      // [0] emit count, [1] "I am the last CTA"
    off += 16;
    s.bar_off = off;
    off += stages * 16;
    const uint32_t replay = ref_replay_smem(k, num_chunks);   // aliases everything above (the scan is over by then)
    s.total = off > replay ? off : replay;
    return s;
}

// Sort n (<= cap, a power of two >= 32) staged entries ascending by the 256 consumer threads and write
// the chunk's segment.  Entries are row << 16 | distance: ascending = row order.
__device__ __forceinline__ void ref_flush_chunk(uint64_t* emit, uint32_t* misc, uint32_t cap, uint8_t* blk,
                                                const RefParams& p, uint32_t chunk, int tid) {
    consumer_bar();
    uint32_t n = lds32_volatile(&misc[0]);
    if (n > cap) {
        if (tid == 0) {
            atomicExch(p.ctrl + 4 * blockIdx.y + kRefCtrlOverflow, 1u);
            atomicExch(p.any_overflow, 1u);
        }
        n = cap;
    }
    uint32_t S = 32;
    while (S < n) S <<= 1;
    for (uint32_t i = n + tid; i < S; i += kConsumerThreads) emit[i] = kNone;
    if (S <= 64) {
        consumer_bar();
        if (tid < 32) {
            volatile uint64_t* v = emit;
            for (uint32_t size = 2; size <= S; size <<= 1)
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t i = tid; i < (S >> 1); i += 32) {
                        const uint32_t pos = 2 * i - (i & (stride - 1));
                        const uint64_t a = v[pos], b = v[pos + stride];
                        if ((a > b) == ((pos & size) == 0)) {
                            v[pos] = b;
                            v[pos + stride] = a;
                        }
                    }
                    __syncwarp();
                }
        }
        consumer_bar();
    } else {
        for (uint32_t size = 2; size <= S; size <<= 1)
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                consumer_bar();
                for (uint32_t i = tid; i < (S >> 1); i += kConsumerThreads) {
                    const uint32_t pos = 2 * i - (i & (stride - 1));
                    const uint64_t a = emit[pos], b = emit[pos + stride];
                    if ((a > b) == ((pos & size) == 0)) {
                        emit[pos] = b;
                        emit[pos + stride] = a;
                    }
                }
            }
        consumer_bar();
    }
    uint64_t* seg = reinterpret_cast<uint64_t*>(blk + p.off_seg) + (size_t)chunk * p.cap;
    for (uint32_t i = tid; i < n; i += kConsumerThreads) seg[i] = emit[i];
    if (tid == 0) {
        reinterpret_cast<uint32_t*>(blk + p.off_cnt)[chunk] = n;
        sts32_volatile(&misc[0], 0u);
    }
    consumer_bar();
}

// After the CTA's last chunk: the last CTA of the query to get here replays.
__device__ __forceinline__ void ref_finish(const RefParams& p, uint8_t* blk, uint32_t* misc, uint8_t* smem, int tid) {
    uint32_t* ctrl = p.ctrl + 4 * blockIdx.y;
    __threadfence();
    consumer_bar();
    if (tid == 0) {
        const uint32_t prev = atomicAdd(ctrl + kRefCtrlDone, 1u);
        sts32_volatile(&misc[1], prev == gridDim.x - 1 ? 1u : 0u);
    }
    consumer_bar();
    if (lds32_volatile(&misc[1]) == 0u) return;
    __threadfence();
    if (tid == 0) ctrl[kRefCtrlDone] = 0;      // leave the counter clean for the next call
    const uint32_t overflow = *reinterpret_cast<volatile uint32_t*>(ctrl + kRefCtrlOverflow);
    if (overflow) return;   // the caller's fallback produces the result
    consumer_bar();         // everyone has read misc before the replay reuses shared memory
    replay_run(replay_args(p, blockIdx.y), smem, tid, kConsumerThreads, 1);
}

// Producer: streams the tiles of this CTA's chunks (rows [P + chunk * chunk_rows, ...)) through the ring.
__device__ __forceinline__ void ref_producer(const RefParams& p, uint32_t tile_rows, uint32_t row_bytes, uint8_t* stage_base,
                                             uint32_t stage_stride, uint64_t* full, uint64_t* empty) {
    const uint32_t S = p.stages;
    const unsigned long long chunk_rows = (unsigned long long)p.chunk_tiles * tile_rows;
    uint32_t s = 0, ph = 0;
    for (uint32_t chunk = blockIdx.x; chunk < p.num_chunks; chunk += gridDim.x) {
        const unsigned long long row0 = p.P + chunk * chunk_rows;
        unsigned long long row1 = row0 + chunk_rows;
        if (row1 > p.N) row1 = p.N;
        for (unsigned long long r = row0; r < row1; r += tile_rows) {
            mbar_wait(&empty[s], ph ^ 1u);
            const unsigned long long left = row1 - r;
            const uint32_t rows = left < tile_rows ? (uint32_t)left : tile_rows;
            const uint32_t bytes = rows * row_bytes, bulk = bytes & ~15u;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.hashes) + r * row_bytes;
            uint8_t* dst = stage_base + s * stage_stride;
            if (bytes != bulk)
                *reinterpret_cast<volatile uint64_t*>(dst + bulk) = *reinterpret_cast<const uint64_t*>(src + bulk);
            mbar_arrive_expect_tx(&full[s], bulk);
            if (bulk) bulk_g2s(dst, src, bulk, &full[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
        }
    }
}

// =============================================================================================
// R1, width-specialised: W words per signature, R rows per thread in registers (as K2, scan.cuh)
// =============================================================================================
template <int W, int R>
__global__ void __launch_bounds__(kScanThreads, 1) ref_scan_kernel(const RefParams p) {
    constexpr int TILE_ROWS = kConsumerThreads * R;
    constexpr uint32_t ROW_BYTES = W * 8;
    constexpr uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;
    constexpr bool EVEN = (W % 2) == 0;
    constexpr int V = W / 2;
    constexpr int VQ = (W + 1) / 2;
    constexpr int G = EVEN ? gcd_int(V, 8) : 1;

    extern __shared__ __align__(128) uint8_t smem[];
    const RefSmem L = ref_smem_layout(TILE_BYTES, p.stages, W, p.cap, p.k, p.num_chunks);
    uint8_t* stage_base = smem + L.stage_off;
    const uint32_t stage_stride = (TILE_BYTES + 127u) & ~127u;
    uint8_t* q_s = smem + L.q_off;
    uint64_t* emit = reinterpret_cast<uint64_t*>(smem + L.emit_off);
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + L.misc_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty = full + p.stages;
    uint8_t* blk = p.ws + (size_t)blockIdx.y * p.ws_stride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
        misc[0] = 0;
        misc[1] = 0;
    }
    // the query: VQ 16-byte vectors + 8 wrap-around copies (bank rotation, as in K2)
    if (tid < VQ + 8) {
        const uint32_t vv = tid < VQ ? tid : (tid - VQ) % VQ;
        const uint64_t* src = p.queries + (size_t)blockIdx.y * W;
        ulonglong2 val;
        val.x = src[2 * vv];
        val.y = (2 * vv + 1 < W) ? src[2 * vv + 1] : 0ull;
        *reinterpret_cast<ulonglong2*>(q_s + tid * 16) = val;
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0) ref_producer(p, TILE_ROWS, ROW_BYTES, stage_base, stage_stride, full, empty);
        return;
    }
    const uint32_t tau = p.ctrl[4 * blockIdx.y + kRefCtrlTau];
    const uint32_t S = p.stages, cap = p.cap;
    const int rot = EVEN ? (((tid & 7) * G) >> 3) : 0;
    const uint32_t qaddr = smem_u32(q_s) + rot * 16;
    const unsigned long long chunk_rows = (unsigned long long)p.chunk_tiles * TILE_ROWS;
    uint32_t s = 0, ph = 0;

    for (uint32_t chunk = blockIdx.x; chunk < p.num_chunks; chunk += gridDim.x) {
        const unsigned long long row0 = p.P + chunk * chunk_rows;
        unsigned long long row1 = row0 + chunk_rows;
        if (row1 > p.N) row1 = p.N;
        for (unsigned long long r0 = row0; r0 < row1; r0 += TILE_ROWS) {
            mbar_wait(&full[s], ph);
            const uint32_t st = smem_u32(stage_base + s * stage_stride);
            uint32_t h[R][2 * W];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t row = r * kConsumerThreads + tid;
                if constexpr (EVEN) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        int jj = v + rot;
                        if (jj >= V) jj -= V;
                        const uint4 x = lds128(st + (row * V + jj) * 16);
                        h[r][4 * v + 0] = x.x;
                        h[r][4 * v + 1] = x.y;
                        h[r][4 * v + 2] = x.z;
                        h[r][4 * v + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const uint2 x = lds64(st + (row * W + w) * 8);
                        h[r][2 * w + 0] = x.x;
                        h[r][2 * w + 1] = x.y;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
            uint32_t d[R];
#pragma unroll
            for (int r = 0; r < R; ++r) d[r] = 0;
#pragma unroll
            for (int v = 0; v < VQ; ++v) {
                const uint4 qv = lds128(qaddr + v * 16);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    d[r] += __popc(h[r][4 * v + 0] ^ qv.x) + __popc(h[r][4 * v + 1] ^ qv.y);
                    if (4 * v + 2 < 2 * W) d[r] += __popc(h[r][4 * v + 2] ^ qv.z) + __popc(h[r][4 * v + 3] ^ qv.w);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const unsigned long long row = r0 + (uint32_t)(r * kConsumerThreads + tid);
                if (d[r] < tau && row < row1) {
                    const uint32_t pos = atomicAdd(&misc[0], 1u);
                    if (pos < cap) emit[pos] = (row << 16) | d[r];
                }
            }
        }
        ref_flush_chunk(emit, misc, cap, blk, p, chunk, tid);
    }
    ref_finish(p, blk, misc, smem, tid);
}

#ifdef NPS_SCAN_GENERIC
// =============================================================================================
// R1, generic width (32 < W <= 63, e.g. the reference's 2560-bit signatures): rows stay in shared
// memory, one row per thread, runtime word loop.
// =============================================================================================
__global__ void __launch_bounds__(kScanThreads, 1) ref_scan_generic_kernel(const RefParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t W = p.words, ROW_BYTES = W * 8, TILE_ROWS = p.tile_rows;
    const uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;
    const RefSmem L = ref_smem_layout(TILE_BYTES, p.stages, W, p.cap, p.k, p.num_chunks);
    uint8_t* stage_base = smem + L.stage_off;
    const uint32_t stage_stride = (TILE_BYTES + 127u) & ~127u;
    uint64_t* q_s = reinterpret_cast<uint64_t*>(smem + L.q_off);
    uint64_t* emit = reinterpret_cast<uint64_t*>(smem + L.emit_off);
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + L.misc_off);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* empty = full + p.stages;
    uint8_t* blk = p.ws + (size_t)blockIdx.y * p.ws_stride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
        misc[0] = 0;
        misc[1] = 0;
    }
    for (uint32_t w = tid; w < W; w += kScanThreads) q_s[w] = p.queries[(size_t)blockIdx.y * W + w];
    __syncthreads();
    if (warp == kConsumerWarps) {
        if (lane == 0) ref_producer(p, TILE_ROWS, ROW_BYTES, stage_base, stage_stride, full, empty);
        return;
    }
    const uint32_t tau = p.ctrl[4 * blockIdx.y + kRefCtrlTau];
    const uint32_t S = p.stages, cap = p.cap;
    const unsigned long long chunk_rows = (unsigned long long)p.chunk_tiles * TILE_ROWS;
    uint32_t s = 0, ph = 0;
    for (uint32_t chunk = blockIdx.x; chunk < p.num_chunks; chunk += gridDim.x) {
        const unsigned long long row0 = p.P + chunk * chunk_rows;
        unsigned long long row1 = row0 + chunk_rows;
        if (row1 > p.N) row1 = p.N;
        for (unsigned long long r0 = row0; r0 < row1; r0 += TILE_ROWS) {
            mbar_wait(&full[s], ph);
            const uint64_t* rows = reinterpret_cast<const uint64_t*>(stage_base + s * stage_stride);
            const unsigned long long row = r0 + (uint32_t)tid;
            const bool valid = (uint32_t)tid < TILE_ROWS && row < row1;
            const uint64_t* my = rows + (size_t)((uint32_t)tid < TILE_ROWS ? tid : 0) * W;
            uint32_t d = 0;
            for (uint32_t w = 0; w < W; ++w) d += __popcll(my[w] ^ q_s[w]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) {
                s = 0;
                ph ^= 1u;
            }
            if (valid && d < tau) {
                const uint32_t pos = atomicAdd(&misc[0], 1u);
                if (pos < cap) emit[pos] = (row << 16) | d;
            }
        }
        ref_flush_chunk(emit, misc, cap, blk, p, chunk, tid);
    }
    ref_finish(p, blk, misc, smem, tid);
}
#endif  // NPS_SCAN_GENERIC

#endif  // NPS_REF_DEVICE_CODE

// host-side launcher table (scan_inst.cu)
typedef cudaError_t (*RefLaunchFn)(const RefParams&, dim3 grid, size_t smem, cudaStream_t);
RefLaunchFn ref_launcher_for_width(int words);   // nullptr if no specialisation (words > 32)
cudaError_t launch_ref_scan_generic(const RefParams&, dim3 grid, size_t smem, cudaStream_t);

// ---- host: plan + run (ref_scan.cu) -------------------------------------------------------------
struct RefPlan {
    bool applicable;
    uint32_t P, tile_rows, chunk_tiles, num_chunks, cap, stages, grid, smem_scan, smem_prefix;
    uint32_t off_pre, off_cnt, off_seg;
    size_t ws_stride;       // bytes per query
};
RefPlan ref_plan(unsigned long long N, uint32_t W, uint32_t k, int sm_count, int smem_optin);
// Workspace: [ref_zero_bytes(nq): overflow word, control words, histograms | nq * plan.ws_stride], 256-byte
// aligned.  The first ref_zero_bytes(nq) bytes must be ZERO when run_ref_top_n is enqueued (zero_first
// does it with one cudaMemsetAsync; a caller that owns the workspace zeroes it once, every call leaves it
// clean).  The uint32 at ws[0] is non-zero afterwards iff some query's segment overflowed (results are then
// not final: the caller runs K5 + the dense replay, gated on that word); it is reset by the next call.
inline size_t ref_zero_bytes(uint32_t nq) {
    return (((size_t)16 + (size_t)nq * 16 + 255) & ~(size_t)255) + (size_t)nq * kRefBins * 4;
}
cudaError_t run_ref_top_n(const RefPlan& pl, const uint64_t* d_hashes, unsigned long long N, uint32_t W,
                          const uint64_t* d_queries, uint32_t nq, uint32_t k, uint64_t* d_rows, uint64_t* d_scores,
                          void* ws, bool zero_first, cudaStream_t stream);

}  // namespace nps
