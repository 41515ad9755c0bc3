// api.cu — the extern "C" surface declared in include/npsims_b200.h: argument checks, launch
// planning, workspace carving.  No torch / Python types cross this boundary.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/npsims_b200.h"
#include "common.cuh"
#include "hamming.cuh"
#include "hist_select.cuh"
#include "merge.cuh"
#include "mma_scan.cuh"
#include "project.cuh"
#include "ref_scan.cuh"
#include "replay.cuh"
#include "rerank.cuh"
#include "scan.cuh"

namespace nps {

unsigned long long g_launch_count = 0;
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define NPS_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) return fail(NPS_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__));       \
    } while (0)

struct DeviceInfo {
    int sm_count = 0;
    int smem_optin = 0;
};

struct ScanPlanInfo {   // mirrors nps_scan_plan_t
    uint32_t tile_rows, tile_bytes, stages, q_tile, num_qtiles, chunk_rows, num_chunks, cap, smem_bytes, grid,
        specialised, rows_per_thread;
};

static int device_info(DeviceInfo* out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    static bool have[64] = {};
    int dev = 0;
    NPS_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(NPS_ECUDA, "device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        NPS_CUDA(cudaDeviceGetAttribute(&cache[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
        NPS_CUDA(cudaDeviceGetAttribute(&cache[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        have[dev] = true;
    }
    *out = cache[dev];
    return NPS_OK;
}

// ---------------------------------------------------------------------------- profiling aid
// When enabled (bench.py's roofline pass only), run_top_n brackets K2 and K3 with CUDA events
// on the launching stream so the dominant kernel can be timed per launch, on the device.
// All of this is PER HOST THREAD (thread_local): the measurement switches, the forced kernel
// family and the "last plan" diagnostics of one thread never affect a call made by another, so the
// library can be driven from many host threads (the reference's ThreadPoolExecutor pattern).  The
// profiling events belong to the device that was current when they were created and are re-created
// when the thread moves to another device.
static thread_local bool g_profile = false;
static thread_local int g_scan_path = NPS_SCAN_AUTO;     // nps_set_scan_path
static thread_local int g_last_path = NPS_SCAN_POPC;     // which kernel family the last top-n call used
static thread_local MmaPlan g_last_mma_plan = {};
static thread_local cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};
static thread_local cudaEvent_t g_phase_ev[2 * kMmaMaxPhases] = {};   // tensor-core path: one pair per phase
static thread_local int g_ev_device = -1;
static thread_local ScanPlanInfo g_last_plan = {};

static int profile_events() {
    int dev = 0;
    NPS_CUDA(cudaGetDevice(&dev));
    if (g_ev_device != dev) {
        for (auto& e : g_ev)
            if (e) {
                cudaEventDestroy(e);
                e = nullptr;
            }
        for (auto& e : g_phase_ev)
            if (e) {
                cudaEventDestroy(e);
                e = nullptr;
            }
        g_ev_device = dev;
    }
    for (auto& e : g_ev)
        if (!e) NPS_CUDA(cudaEventCreate(&e));
    for (auto& e : g_phase_ev)
        if (!e) NPS_CUDA(cudaEventCreate(&e));
    return NPS_OK;
}

struct DeviceGuard {   // host-level entry points leave the caller's current device as it was
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

// ---------------------------------------------------------------------------- scan planning
struct ScanPlan {
    bool specialised;
    uint32_t tile_rows, tile_bytes, stages, q_tile, num_qtiles, chunk_rows, num_chunks, cap, row_bits;
    uint32_t smem, lists_per_q;
    int grid;
};

static uint32_t bit_length(uint64_t v) {
    uint32_t b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return b;
}

// Pure function of the shape and the device: the workspace query and the launch agree.
static int plan_scan(uint64_t N, uint32_t W, uint32_t Q, uint32_t k, int sm_count, int smem_optin, ScanPlan* pl) {
    if (W == 0 || W > NPS_MAX_WORDS) return fail(NPS_EINVAL, "words=%u outside [1, %u]", W, NPS_MAX_WORDS);
    if (k == 0 || k > NPS_MAX_K) return fail(NPS_EINVAL, "k=%u outside [1, %u]", k, NPS_MAX_K);
    if (Q == 0) return fail(NPS_EINVAL, "num_queries must be > 0");
    if (sm_count <= 0) sm_count = 148;
    if (smem_optin <= 0) smem_optin = 232448;
    pl->specialised = W <= 32;
    const uint32_t dist_bits = bit_length(64ull * W + 1);   // max distance 64W < 2^dist_bits - 1
    pl->row_bits = 32 - dist_bits;
    uint32_t cap = 64;
    while (cap < k + 32) cap <<= 1;
    pl->cap = cap;
    if (pl->specialised) {
        pl->tile_rows = kConsumerThreads * rows_per_thread((int)W);
    } else {
        uint32_t rows = (48u * 1024u) / (W * 8u);
        rows &= ~1u;
        if (rows > (uint32_t)kConsumerThreads) rows = kConsumerThreads;
        if (rows < 2) return fail(NPS_EINVAL, "words=%u too wide", W);
        pl->tile_rows = rows;
    }
    pl->tile_bytes = pl->tile_rows * W * 8;
    const uint32_t stride = (pl->tile_bytes + 127u) & ~127u;
    const uint32_t per_q = scan_q_stride(W) + cap * 4 + 12;
    const uint32_t want_q = Q < 128 ? Q : 128;
    uint32_t stages = 0, q_tile_max = 0;
    for (uint32_t S = 4; S >= 2; --S) {
        const int64_t avail = (int64_t)smem_optin - (int64_t)S * stride - (int64_t)S * 16 - 64;
        if (avail < (int64_t)per_q) continue;
        const uint32_t qmax = (uint32_t)(avail / per_q);
        if (qmax >= (want_q < 32 ? want_q : 32) || S == 2) {
            stages = S;
            q_tile_max = qmax;
            break;
        }
    }
    if (stages == 0 || q_tile_max == 0)
        return fail(NPS_EINVAL, "k=%u with words=%u does not fit in shared memory", k, W);
    pl->stages = stages;
    uint32_t q_tile = want_q < q_tile_max ? want_q : q_tile_max;
    pl->num_qtiles = (Q + q_tile - 1) / q_tile;
    pl->q_tile = (Q + pl->num_qtiles - 1) / pl->num_qtiles;

    const uint64_t tiles_total = N == 0 ? 1 : (N + pl->tile_rows - 1) / pl->tile_rows;
    uint64_t max_tpc = (1ull << pl->row_bits) / pl->tile_rows;
    if (max_tpc < 1) return fail(NPS_EINVAL, "words=%u: chunk does not fit the key", W);
    if (max_tpc > tiles_total) max_tpc = tiles_total;
    // choose tiles-per-chunk minimising  waves * (tiles + fixed cost)  + merge cost
    double best_cost = 1e300;
    uint64_t best_tpc = max_tpc;
    const double merge_w = 0.002 * (double)(k < 64 ? 64 : k) / 64.0;
    auto consider = [&](uint64_t tpc) {
        if (tpc < 1 || tpc > max_tpc) return;
        const uint64_t chunks = (tiles_total + tpc - 1) / tpc;
        const uint64_t items = chunks * pl->num_qtiles;
        const uint64_t waves = (items + sm_count - 1) / sm_count;
        const double cost = (double)waves * ((double)tpc + 0.5) + merge_w * (double)chunks;
        if (cost < best_cost) {
            best_cost = cost;
            best_tpc = tpc;
        }
    };
    for (uint64_t tpc = 1; tpc <= max_tpc && tpc <= 4096; ++tpc) consider(tpc);
    for (uint64_t c = 1; c <= (uint64_t)sm_count * 8; ++c) consider((tiles_total + c - 1) / c);
    consider(max_tpc);
    pl->chunk_rows = (uint32_t)(best_tpc * pl->tile_rows);
    pl->num_chunks = (uint32_t)((tiles_total + best_tpc - 1) / best_tpc);
    const uint64_t items = (uint64_t)pl->num_chunks * pl->num_qtiles;
    if (items > 0xFFFFFFFFull) return fail(NPS_EINVAL, "too many work items");
    pl->grid = (int)(items < (uint64_t)sm_count ? items : (uint64_t)sm_count);
    // One or two queries (the single-query scan): one shared list per query means eight consumer warps
    // on one lock; give every warp its own list (merged at the end of the work item).  From ~4 queries
    // on the eight-fold appends of the private lists cost more than the lock (measured).
    pl->lists_per_q = (pl->specialised && pl->q_tile <= 2 && cap == 64 && k <= 32) ? (uint32_t)kConsumerWarps : 1u;
    pl->smem = scan_smem_layout(pl->tile_bytes, pl->stages, W, pl->q_tile, cap, pl->lists_per_q).total;
    if ((int)pl->smem > smem_optin) return fail(NPS_EINVAL, "internal: smem plan %u > %d", pl->smem, smem_optin);
    return NPS_OK;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Batched queries go to the tensor cores (mma_scan.cu); few queries stay on the HBM-bound
// XOR+POPC streaming scan (scan.cuh).  NPS_SCAN_POPC / NPS_SCAN_MMA force one family.
static bool choose_mma(uint64_t N, uint32_t W, uint32_t Q, uint32_t k, const DeviceInfo& dev, MmaPlan* mp) {
    if (g_scan_path == NPS_SCAN_POPC) return false;
    // AUTO (measured, profiles/r01_c5_sweep.json): the tensor-core phases have ~0.1 ms of fixed cost
    // (7 scan + 7 select launches); they win from 64 queries on any table worth batching, and from
    // 16 queries once the batch holds >= 1e9 (row, query, word) triples.
    if (g_scan_path == NPS_SCAN_AUTO &&
        !((Q >= 64 && N >= 4096) || (Q >= 16 && (double)N * Q * W >= 1e9)))
        return false;
    return mma_plan(N, W, Q, k, dev.sm_count, dev.smem_optin, mp);
}

// Few queries (<= 32): counting select on the distance histogram (hist_select.cuh) instead of K2's sorted
// candidate lists — any k at the cost of the table stream (measured, 16M rows x 640 bit, 1 query:
// k = 10 / 100 / 1000 in 0.27 ms each; K2: 0.31 / 0.82 / 2.91 ms).  K2 + K3 remain for signatures wider
// than 63 words, tiny tables, larger POPC batches and the gated overflow repair of the tensor-core path.
static bool choose_hist(uint64_t N, uint32_t W, uint32_t Q, uint32_t k, const DeviceInfo& dev, HistPlan* hp) {
    if (W > kHistMaxWords || Q > 32 || N < 4096) return false;
    *hp = hist_plan(N, W, Q, k, dev.sm_count, dev.smem_optin);
    return hp->applicable;
}

struct TopNWorkspace {
    size_t keys_bytes, scratch_bytes, total;
};
static TopNWorkspace topn_workspace(const ScanPlan& pl, uint32_t Q, uint32_t k) {
    TopNWorkspace w;
    w.keys_bytes = align256((size_t)Q * pl.num_chunks * k * 8);
    w.scratch_bytes = align256(merge_scratch_keys(Q, pl.num_chunks, k) * 8);
    w.total = w.keys_bytes + w.scratch_bytes;
    return w;
}

// The XOR+POPC scan (K2) + block merge (K3) over one batch of queries.  `gate` (device pointer or
// null): when non-null every launch exits at once unless *gate != 0 — that is how the tensor-core
// path repairs a candidate-buffer overflow on the device, with no host round trip.
static int run_popc(const ScanPlan& pl, const uint64_t* d_hashes, uint64_t N, uint32_t W, const uint64_t* d_queries,
                    uint32_t Q, uint32_t k, uint64_t row_base, uint64_t* d_rows, uint64_t* d_dists,
                    uint64_t* d_keys_out, void* ws, const uint32_t* gate, bool profile, cudaStream_t stream) {
    const TopNWorkspace w = topn_workspace(pl, Q, k);
    uint64_t* keys = reinterpret_cast<uint64_t*>(ws);
    uint64_t* scratch = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ws) + w.keys_bytes);

    ScanParams p;
    p.hashes = d_hashes;
    p.queries = d_queries;
    p.out_keys = keys;
    p.gate = gate;
    p.N = N;
    p.row_base = row_base;
    p.Q = Q;
    p.k = k;
    p.cap = pl.cap;
    p.words = W;
    p.chunk_rows = pl.chunk_rows;
    p.num_chunks = pl.num_chunks;
    p.q_tile = pl.q_tile;
    p.num_qtiles = pl.num_qtiles;
    p.row_bits = pl.row_bits;
    p.stages = pl.stages;
    p.tile_rows = pl.tile_rows;
    p.lists_per_q = pl.lists_per_q;
    if (profile) NPS_CUDA(cudaEventRecord(g_ev[0], stream));
    if (pl.specialised) {
        ScanLaunchFn fn = scan_launcher_for_width((int)W);
        if (!fn) return fail(NPS_EINVAL, "no scan kernel for words=%u", W);
        NPS_CUDA(fn(p, pl.grid, pl.smem, stream));
    } else {
        NPS_CUDA(launch_scan_generic(p, pl.grid, pl.smem, stream));
    }
    if (profile) NPS_CUDA(cudaEventRecord(g_ev[1], stream));
    NPS_CUDA(launch_merge(keys, Q, pl.num_chunks, k, /*list_major=*/0, scratch, d_rows, d_dists, d_keys_out, stream,
                          gate));
    if (profile) NPS_CUDA(cudaEventRecord(g_ev[2], stream));
    return NPS_OK;
}

// Device-side overflow repair of the tensor-core path: the same batch on K2/K3, in query slices
// whose workspace stays under kRepairSliceBytes, every launch gated on the overflow counter.
constexpr size_t kRepairSliceBytes = 256ull << 20;
struct RepairPlan {
    uint32_t slice_q;     // queries per slice
    ScanPlan full, tail;  // plans of a full slice and of the last (shorter) one
    size_t ws_bytes;
};
static int plan_repair(uint64_t N, uint32_t W, uint32_t Q, uint32_t k, const DeviceInfo& dev, RepairPlan* rp) {
    uint32_t sq = Q;
    for (;;) {
        int rc = plan_scan(N, W, sq, k, dev.sm_count, dev.smem_optin, &rp->full);
        if (rc) return rc;
        if (topn_workspace(rp->full, sq, k).total <= kRepairSliceBytes || sq <= 64) break;
        sq = (sq + 1) / 2;
    }
    rp->slice_q = sq;
    rp->ws_bytes = topn_workspace(rp->full, sq, k).total;
    rp->tail = rp->full;
    if (Q % sq) {
        int rc = plan_scan(N, W, Q % sq, k, dev.sm_count, dev.smem_optin, &rp->tail);
        if (rc) return rc;
        const size_t t = topn_workspace(rp->tail, Q % sq, k).total;
        if (t > rp->ws_bytes) rp->ws_bytes = t;
    }
    return NPS_OK;
}

static int run_top_n(const uint64_t* d_hashes, uint64_t N, uint32_t W, const uint64_t* d_queries, uint32_t Q,
                     uint32_t k, uint64_t row_base, uint64_t* d_rows, uint64_t* d_dists, uint64_t* d_keys_out,
                     void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!d_hashes && N > 0) return fail(NPS_EINVAL, "d_hashes is null");
    if (!d_queries) return fail(NPS_EINVAL, "d_queries is null");
    if (!d_rows && !d_keys_out) return fail(NPS_EINVAL, "output pointer is null");
    if (reinterpret_cast<uintptr_t>(d_hashes) & 15u) return fail(NPS_EINVAL, "d_hashes must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_queries) & 7u) return fail(NPS_EINVAL, "d_queries must be 8-byte aligned");
    if (N + row_base >= (1ull << kGlobalRowBits)) return fail(NPS_EINVAL, "row ids exceed 2^40");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    if (W == 0 || W > NPS_MAX_WORDS) return fail(NPS_EINVAL, "words=%u outside [1, %u]", W, NPS_MAX_WORDS);
    if (k == 0 || k > NPS_MAX_K) return fail(NPS_EINVAL, "k=%u outside [1, %u]", k, NPS_MAX_K);
    if (Q == 0) return fail(NPS_EINVAL, "num_queries must be > 0");
    MmaPlan mp;
    if (choose_mma(N, W, Q, k, dev, &mp)) {
        RepairPlan rp;
        rc = plan_repair(N, W, Q, k, dev, &rp);
        if (rc) return rc;
        const size_t need = align256(mp.total) + rp.ws_bytes;
        if (ws_bytes < need || !ws) return fail(NPS_EWORKSPACE, "workspace %zu bytes < required %zu", ws_bytes, need);
        if (reinterpret_cast<uintptr_t>(ws) & 255u) return fail(NPS_EINVAL, "workspace must be 256-byte aligned");
        g_last_path = NPS_SCAN_MMA;
        g_last_mma_plan = mp;
        if (g_profile) {
            rc = profile_events();
            if (rc) return rc;
            NPS_CUDA(cudaEventRecord(g_ev[0], stream));
        }
        NPS_CUDA(run_mma_top_n(mp, d_hashes, N, W, d_queries, Q, k, row_base, d_rows, d_dists, d_keys_out, ws, stream,
                               g_profile ? g_phase_ev : nullptr));
        if (g_profile) NPS_CUDA(cudaEventRecord(g_ev[1], stream));   // scan = all phases (scan + select launches)
        // repair: the overflow counter is the first word of the workspace (mma_scan.cu)
        const uint32_t* gate = reinterpret_cast<const uint32_t*>(ws);
        void* rws = reinterpret_cast<uint8_t*>(ws) + align256(mp.total);
        for (uint32_t q0 = 0; q0 < Q; q0 += rp.slice_q) {
            const uint32_t nq = Q - q0 < rp.slice_q ? Q - q0 : rp.slice_q;
            const size_t o = (size_t)q0 * k;
            rc = run_popc(nq == rp.slice_q ? rp.full : rp.tail, d_hashes, N, W, d_queries + (size_t)q0 * W, nq, k,
                          row_base, d_rows ? d_rows + o : nullptr, d_dists ? d_dists + o : nullptr,
                          d_keys_out ? d_keys_out + o : nullptr, rws, gate, false, stream);
            if (rc) return rc;
        }
        if (g_profile) NPS_CUDA(cudaEventRecord(g_ev[2], stream));   // "merge" slot = the gated repair launches
        return NPS_OK;
    }
    g_last_path = NPS_SCAN_POPC;
    HistPlan hp;
    if (choose_hist(N, W, Q, k, dev, &hp)) {
        if (ws_bytes < hp.total || !ws) return fail(NPS_EWORKSPACE, "workspace %zu bytes < required %zu", ws_bytes, hp.total);
        if (reinterpret_cast<uintptr_t>(ws) & 255u) return fail(NPS_EINVAL, "workspace must be 256-byte aligned");
        g_last_plan = ScanPlanInfo{hp.tile_rows, hp.tile_rows * W * 8, hp.stages, hp.q_tile, hp.num_qtiles, hp.chunk_rows,
                                   hp.num_chunks, 0, hp.smem, hp.grid, W <= 32 ? 1u : 0u,
                                   W <= 32 ? (uint32_t)rows_per_thread((int)W) : 1u};
        if (g_profile) {
            rc = profile_events();
            if (rc) return rc;
        }
        NPS_CUDA(run_hist_top_n(hp, d_hashes, N, W, d_queries, Q, k, row_base, d_rows, d_dists, d_keys_out, ws, stream,
                                g_profile ? g_ev : nullptr));
        return NPS_OK;
    }
    ScanPlan pl;
    rc = plan_scan(N, W, Q, k, dev.sm_count, dev.smem_optin, &pl);
    if (rc) return rc;
    const TopNWorkspace w = topn_workspace(pl, Q, k);
    if (ws_bytes < w.total || (!ws && w.total))
        return fail(NPS_EWORKSPACE, "workspace %zu bytes < required %zu", ws_bytes, w.total);
    if (reinterpret_cast<uintptr_t>(ws) & 15u) return fail(NPS_EINVAL, "workspace must be 16-byte aligned");
    g_last_plan = ScanPlanInfo{pl.tile_rows, pl.tile_bytes, pl.stages, pl.q_tile, pl.num_qtiles, pl.chunk_rows,
                               pl.num_chunks, pl.cap, pl.smem, (uint32_t)pl.grid, pl.specialised ? 1u : 0u,
                               pl.specialised ? (uint32_t)rows_per_thread((int)W) : 1u};
    if (g_profile) {
        rc = profile_events();
        if (rc) return rc;
    }
    return run_popc(pl, d_hashes, N, W, d_queries, Q, k, row_base, d_rows, d_dists, d_keys_out, ws, nullptr, g_profile,
                    stream);
}

}  // namespace nps

using namespace nps;

extern "C" {

int nps_version(void) { return 100; }
const char* nps_last_error(void) { return g_err; }

int nps_profile_enable(int on) {
    g_profile = on != 0;
    return NPS_OK;
}

int nps_profile_last(float* scan_ms, float* merge_ms) {
    if (!g_ev[0] || !g_ev[2]) return fail(NPS_EINVAL, "no profiled call yet");
    NPS_CUDA(cudaEventSynchronize(g_ev[2]));
    if (scan_ms) NPS_CUDA(cudaEventElapsedTime(scan_ms, g_ev[0], g_ev[1]));
    if (merge_ms) NPS_CUDA(cudaEventElapsedTime(merge_ms, g_ev[1], g_ev[2]));
    return NPS_OK;
}

int nps_last_scan_plan(nps_scan_plan_t* out) {
    if (!out) return fail(NPS_EINVAL, "out is null");
    static_assert(sizeof(nps_scan_plan_t) == sizeof(ScanPlanInfo), "plan struct mismatch");
    memcpy(out, &g_last_plan, sizeof(*out));
    return NPS_OK;
}

int nps_set_scan_path(int path) {
    if (path != NPS_SCAN_AUTO && path != NPS_SCAN_POPC && path != NPS_SCAN_MMA) return fail(NPS_EINVAL, "bad scan path %d", path);
    g_scan_path = path;
    return NPS_OK;
}

int nps_last_scan_path(void) { return g_last_path; }

int nps_set_mma_pair(int mode) {
    if (mode < 0 || mode > 3) return fail(NPS_EINVAL, "bad pair mode %d", mode);
    mma_set_pair_mode(mode);
    return NPS_OK;
}

int nps_last_mma_plan(nps_mma_plan_t* out) {
    if (!out) return fail(NPS_EINVAL, "out is null");
    const MmaPlan& m = g_last_mma_plan;
    const MmaPhase& L = m.phase[m.num_phases ? m.num_phases - 1 : 0];
    *out = nps_mma_plan_t{m.n_tile, m.num_qtiles, m.cap, m.growth, m.a_stages, m.smem, m.num_phases,
                          L.tiles, L.tiles_per_chunk, L.num_chunks, L.grid, m.pair | (m.qt << 8)};
    return NPS_OK;
}

int nps_profile_last_phases(float* scan_ms, uint32_t* tiles, uint32_t max_phases, uint32_t* num_phases) {
    if (!scan_ms || !tiles || !num_phases) return fail(NPS_EINVAL, "null pointer");
    if (g_last_path != NPS_SCAN_MMA || !g_phase_ev[0]) return fail(NPS_EINVAL, "no profiled tensor-core call yet");
    const MmaPlan& m = g_last_mma_plan;
    NPS_CUDA(cudaEventSynchronize(g_ev[2]));
    uint32_t n = 0;
    for (uint32_t ph = 0; ph < m.num_phases && n < max_phases; ++ph, ++n) {
        tiles[n] = m.phase[ph].tiles;
        scan_ms[n] = 0.f;
        if (m.phase[ph].tiles) NPS_CUDA(cudaEventElapsedTime(&scan_ms[n], g_phase_ev[2 * ph], g_phase_ev[2 * ph + 1]));
    }
    *num_phases = n;
    return NPS_OK;
}

int nps_top_n_overflow_count(const void* d_workspace, void* stream, uint32_t* out_count) {
    if (!d_workspace || !out_count) return fail(NPS_EINVAL, "null pointer");
    NPS_CUDA(cudaMemcpyAsync(out_count, d_workspace, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    NPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return NPS_OK;
}

unsigned long long nps_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }

int nps_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(NPS_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int nps_hamming(const uint64_t* d_hashes, uint64_t num_hashes, uint32_t words, const uint64_t* d_queries,
                uint32_t num_queries, void* d_out, int out_bytes, void* stream) {
    if (words == 0 || words > NPS_MAX_WORDS) return fail(NPS_EINVAL, "words=%u outside [1, %u]", words, NPS_MAX_WORDS);
    if (out_bytes != 2 && out_bytes != 4 && out_bytes != 8) return fail(NPS_EINVAL, "out_bytes must be 2, 4 or 8");
    if (out_bytes == 2 && 64ull * words > 65535ull) return fail(NPS_EINVAL, "uint16 output overflows for words=%u", words);
    if (num_hashes == 0 || num_queries == 0) return NPS_OK;
    if (!d_hashes || !d_queries || !d_out) return fail(NPS_EINVAL, "null pointer");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    NPS_CUDA(launch_hamming_dist(d_hashes, num_hashes, words, d_queries, num_queries, d_out, out_bytes, dev.sm_count,
                                 (cudaStream_t)stream));
    return NPS_OK;
}

size_t nps_hamming_top_n_workspace_bytes(uint64_t num_hashes, uint32_t words, uint32_t num_queries, uint32_t k) {
    DeviceInfo dev;
    if (device_info(&dev) != NPS_OK) {   // no device (CPU-only host): size for a B200
        dev.sm_count = 148;
        dev.smem_optin = 232448;
    }
    if (words == 0 || words > NPS_MAX_WORDS || k == 0 || k > NPS_MAX_K || num_queries == 0) {
        fail(NPS_EINVAL, "bad shape: words=%u k=%u num_queries=%u", words, k, num_queries);
        return 0;
    }
    MmaPlan mp;
    if (choose_mma(num_hashes, words, num_queries, k, dev, &mp)) {
        RepairPlan rp;
        if (plan_repair(num_hashes, words, num_queries, k, dev, &rp) != NPS_OK) return 0;
        return align256(mp.total) + rp.ws_bytes;
    }
    HistPlan hp;
    if (choose_hist(num_hashes, words, num_queries, k, dev, &hp)) return hp.total;
    ScanPlan pl;
    if (plan_scan(num_hashes, words, num_queries, k, dev.sm_count, dev.smem_optin, &pl) != NPS_OK) return 0;
    return topn_workspace(pl, num_queries, k).total;
}

int nps_hamming_top_n(const uint64_t* d_hashes, uint64_t num_hashes, uint32_t words, const uint64_t* d_queries,
                      uint32_t num_queries, uint32_t k, uint64_t row_base, uint64_t* d_out_rows,
                      uint64_t* d_out_dists, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_out_rows) return fail(NPS_EINVAL, "d_out_rows is null");
    return run_top_n(d_hashes, num_hashes, words, d_queries, num_queries, k, row_base, d_out_rows, d_out_dists,
                     nullptr, d_workspace, workspace_bytes, (cudaStream_t)stream);
}

int nps_hamming_top_n_keys(const uint64_t* d_hashes, uint64_t num_hashes, uint32_t words, const uint64_t* d_queries,
                           uint32_t num_queries, uint32_t k, uint64_t row_base, uint64_t* d_out_keys,
                           void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!d_out_keys) return fail(NPS_EINVAL, "d_out_keys is null");
    return run_top_n(d_hashes, num_hashes, words, d_queries, num_queries, k, row_base, nullptr, nullptr, d_out_keys,
                     d_workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t nps_top_n_merge_workspace_bytes(uint32_t num_queries, uint32_t num_lists, uint32_t k) {
    return align256(merge_scratch_keys(num_queries, num_lists, k) * 8);
}

int nps_top_n_merge(const uint64_t* d_keys, uint32_t num_lists, uint32_t num_queries, uint32_t k, int list_major,
                    uint64_t* d_out_rows, uint64_t* d_out_dists, void* d_workspace, size_t workspace_bytes,
                    void* stream) {
    if (k == 0 || k > NPS_MAX_K) return fail(NPS_EINVAL, "k=%u outside [1, %u]", k, NPS_MAX_K);
    if (num_lists == 0) return fail(NPS_EINVAL, "num_lists must be > 0");
    if (num_queries == 0) return NPS_OK;
    if (!d_keys || !d_out_rows) return fail(NPS_EINVAL, "null pointer");
    const size_t need = nps_top_n_merge_workspace_bytes(num_queries, num_lists, k);
    if (need && (!d_workspace || workspace_bytes < need))
        return fail(NPS_EWORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, need);
    NPS_CUDA(launch_merge(d_keys, num_queries, num_lists, k, list_major, reinterpret_cast<uint64_t*>(d_workspace),
                          d_out_rows, d_out_dists, nullptr, (cudaStream_t)stream));
    return NPS_OK;
}

// Workspace of the reference-order search: [header | nq x K2R blocks | nq x dense distance rows (fallback)]
struct RefWorkspace {
    RefPlan plan;
    size_t fast_bytes, row_bytes, total;
    int elt;
};
static RefWorkspace ref_workspace(uint64_t N, uint32_t W, uint32_t nq, uint32_t k, const DeviceInfo& dev) {
    RefWorkspace w;
    w.plan = ref_plan(N, W, k, dev.sm_count, dev.smem_optin);
    w.elt = (64ull * W > 65535ull) ? 4 : 2;
    w.row_bytes = ((size_t)N * w.elt + 15) & ~(size_t)15;      // every query's distance row starts 16-byte aligned
    w.fast_bytes = w.plan.applicable ? align256(ref_zero_bytes(nq) + (size_t)nq * w.plan.ws_stride) : 0;
    w.total = w.fast_bytes + align256(w.row_bytes * nq);
    return w;
}

size_t nps_hamming_top_n_reference_workspace_bytes(uint64_t num_hashes, uint32_t words, uint32_t num_queries, uint32_t k) {
    DeviceInfo dev;
    if (device_info(&dev) != NPS_OK) {
        dev.sm_count = 148;
        dev.smem_optin = 232448;
    }
    if (words == 0 || words > NPS_MAX_WORDS || k == 0 || k > NPS_MAX_K) return 0;
    return ref_workspace(num_hashes, words, num_queries ? num_queries : 1, k, dev).total;
}

int nps_hamming_top_n_reference(const uint64_t* d_hashes, uint64_t num_hashes, uint32_t words,
                                const uint64_t* d_queries, uint32_t num_queries, uint32_t k, uint64_t* d_out_rows,
                                uint64_t* d_out_scores, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (words == 0 || words > NPS_MAX_WORDS) return fail(NPS_EINVAL, "words=%u outside [1, %u]", words, NPS_MAX_WORDS);
    if (k == 0 || k > NPS_MAX_K) return fail(NPS_EINVAL, "k=%u outside [1, %u]", k, NPS_MAX_K);
    if (num_queries == 0) return NPS_OK;
    if (!d_queries || !d_out_rows || (!d_hashes && num_hashes)) return fail(NPS_EINVAL, "null pointer");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    const RefWorkspace w = ref_workspace(num_hashes, words, num_queries, k, dev);
    if (!d_workspace || workspace_bytes < w.total)
        return fail(NPS_EWORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, w.total);
    if (reinterpret_cast<uintptr_t>(d_workspace) & 255u) return fail(NPS_EINVAL, "workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    void* dense = reinterpret_cast<uint8_t*>(d_workspace) + w.fast_bytes;
    const uint32_t* gate = nullptr;
    const bool fast = w.plan.applicable && (reinterpret_cast<uintptr_t>(d_hashes) & 15u) == 0;
    if (fast) {
        // K2R: prefix kernel + streaming superset scan + replay by the last CTA (ref_scan.cuh)
        NPS_CUDA(run_ref_top_n(w.plan, d_hashes, num_hashes, words, d_queries, num_queries, k, d_out_rows, d_out_scores,
                               d_workspace, /*zero_first=*/true, s));
        gate = reinterpret_cast<const uint32_t*>(d_workspace);   // fallback below runs only if a segment overflowed
        if (num_hashes <= kRefSmallRows) return NPS_OK;           // all prefix: nothing can overflow
    }
    // K5 + dense replay: the only path for signatures wider than kRefMaxWords words, the gated
    // fallback otherwise.  The replay kernel indexes query q's distances at q * N elements, so rows
    // are packed; it checks alignment per row and falls back to scalar loads.
    NPS_CUDA(launch_hamming_dist(d_hashes, num_hashes, words, d_queries, num_queries, dense, w.elt, dev.sm_count, s, gate));
    NPS_CUDA(launch_replay_queue(dense, w.elt, num_hashes, num_queries, k, d_out_rows, d_out_scores, s, gate));
    return NPS_OK;
}

int nps_num_unshared_bits(const uint64_t* d_a, uint64_t len_a, const uint64_t* d_b, uint64_t len_b, uint64_t n,
                          uint8_t* d_out, void* stream) {
    if (n == 0) return NPS_OK;
    if (!d_a || !d_b || !d_out || len_a == 0 || len_b == 0) return fail(NPS_EINVAL, "null pointer / empty operand");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    NPS_CUDA(launch_unshared_bits(d_a, len_a, d_b, len_b, n, d_out, dev.sm_count, (cudaStream_t)stream));
    return NPS_OK;
}

int nps_rerank_top_n(const uint64_t* d_hashes, uint64_t num_hashes, uint32_t words, const uint64_t* d_queries,
                     uint32_t num_queries, const uint64_t* d_candidates, uint32_t num_candidates, uint32_t k,
                     uint64_t* d_out_rows, uint64_t* d_out_dists, void* stream) {
    if (words == 0 || words > NPS_MAX_WORDS) return fail(NPS_EINVAL, "words=%u outside [1, %u]", words, NPS_MAX_WORDS);
    if (k == 0 || k > NPS_MAX_K) return fail(NPS_EINVAL, "k=%u outside [1, %u]", k, NPS_MAX_K);
    if (num_candidates == 0 || num_candidates > 4096) return fail(NPS_EINVAL, "num_candidates=%u outside [1, 4096]", num_candidates);
    if (num_queries == 0) return NPS_OK;
    if (!d_hashes || !d_queries || !d_candidates || !d_out_rows) return fail(NPS_EINVAL, "null pointer");
    if (reinterpret_cast<uintptr_t>(d_hashes) & 15u) return fail(NPS_EINVAL, "d_hashes must be 16-byte aligned");
    NPS_CUDA(launch_rerank(d_hashes, num_hashes, words, d_queries, num_queries, d_candidates, num_candidates, k,
                           d_out_rows, d_out_dists, (cudaStream_t)stream));
    return NPS_OK;
}

int nps_project_sign_pack_f64(const void* d_vectors, int vectors_are_f32, uint64_t num_vectors, uint32_t dims,
                              const double* d_projections, uint32_t num_projections, uint64_t* d_out, double* d_dots,
                              void* stream) {
    if (num_projections == 0 || num_projections % 64 != 0)
        return fail(NPS_EINVAL, "Projections must be a multiple of 64");   // lsh.py:37-38
    if (dims == 0) return fail(NPS_EINVAL, "dims must be > 0");
    if (num_vectors == 0) return NPS_OK;
    if (!d_vectors || !d_projections || (!d_out && !d_dots)) return fail(NPS_EINVAL, "null pointer");
    NPS_CUDA(launch_project_f64(d_vectors, vectors_are_f32, num_vectors, dims, d_projections, num_projections, d_out,
                                d_dots, (cudaStream_t)stream));
    return NPS_OK;
}

size_t nps_project_sign_pack_workspace_bytes(uint64_t num_vectors, uint32_t num_projections) {
    return project_tc_workspace_bytes(num_vectors, num_projections, nullptr);
}

size_t nps_project_split_bytes(uint32_t num_projections, uint32_t dims) { return project_split_bytes(num_projections, dims); }

static thread_local int g_hash_path = NPS_HASH_AUTO;
int nps_set_hash_path(int path) {
    if (path != NPS_HASH_AUTO && path != NPS_HASH_TF32 && path != NPS_HASH_BF16) return fail(NPS_EINVAL, "bad hash path %d", path);
    g_hash_path = path;
    return NPS_OK;
}

int nps_project_split_projections(const double* d_projections_f64, uint32_t num_projections, uint32_t dims,
                                  float* d_out_split, void* stream) {
    if (!d_projections_f64 || !d_out_split) return fail(NPS_EINVAL, "null pointer");
    NPS_CUDA(launch_split_projections(d_projections_f64, num_projections, dims, d_out_split, (cudaStream_t)stream));
    return NPS_OK;
}

int nps_project_sign_pack(const float* d_vectors, uint64_t num_vectors, uint32_t dims, const float* d_projections_f32,
                          const double* d_projections_f64, uint32_t num_projections, float projection_norm_max,
                          uint64_t* d_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (num_projections == 0 || num_projections % 64 != 0)
        return fail(NPS_EINVAL, "Projections must be a multiple of 64");   // lsh.py:37-38
    if (num_vectors == 0) return NPS_OK;
    if (!d_vectors || !d_projections_f32 || !d_projections_f64 || !d_out || !d_workspace) return fail(NPS_EINVAL, "null pointer");
    if (!project_tc_applicable(num_vectors, dims, num_projections))
        return fail(NPS_EINVAL, "tensor-core projection needs float32 vectors with dims %% 4 == 0 (dims=%u); use "
                                "nps_project_sign_pack_f64", dims);
    if ((reinterpret_cast<uintptr_t>(d_vectors) | reinterpret_cast<uintptr_t>(d_projections_f32)) & 15u)
        return fail(NPS_EINVAL, "vectors and projections must be 16-byte aligned");
    if (workspace_bytes < project_tc_workspace_bytes(num_vectors, num_projections, nullptr))
        return fail(NPS_EWORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes,
                    project_tc_workspace_bytes(num_vectors, num_projections, nullptr));
    if (!(projection_norm_max > 0.f)) return fail(NPS_EINVAL, "projection_norm_max must be > 0");
    DeviceInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    NPS_CUDA(launch_project_tc(d_vectors, num_vectors, dims, d_projections_f32, d_projections_f64, num_projections,
                               projection_norm_max, d_out, d_workspace, dev.sm_count, dev.smem_optin,
                               (cudaStream_t)stream, g_hash_path != NPS_HASH_TF32));
    return NPS_OK;
}

int nps_project_sign_pack_stats(const void* d_workspace, uint64_t num_vectors, uint32_t num_projections, void* stream,
                                uint32_t* recomputed, uint32_t* capacity) {
    if (!d_workspace || !recomputed || !capacity) return fail(NPS_EINVAL, "null pointer");
    uint32_t cap = 0;
    project_tc_workspace_bytes(num_vectors, num_projections, &cap);
    const size_t off = ((size_t)num_vectors * 4 + 255) / 256 * 256;
    NPS_CUDA(cudaMemcpyAsync(recomputed, reinterpret_cast<const uint8_t*>(d_workspace) + off, 4, cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    NPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *capacity = cap;
    return NPS_OK;
}

int nps_project_sign_pack_stats_in_kernel(const void* d_workspace, uint64_t num_vectors, void* stream,
                                          uint32_t* in_kernel) {
    if (!d_workspace || !in_kernel) return fail(NPS_EINVAL, "null pointer");
    const size_t off = ((size_t)num_vectors * 4 + 255) / 256 * 256 + 4;
    NPS_CUDA(cudaMemcpyAsync(in_kernel, reinterpret_cast<const uint8_t*>(d_workspace) + off, 4, cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    NPS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return NPS_OK;
}

// ------------------------------------------------------------------------------ host level
struct nps_index {
    int device = 0;
    uint64_t num_hashes = 0;
    uint32_t words = 0;
    uint64_t* d_hashes = nullptr;
    cudaStream_t stream = nullptr;
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    uint64_t* d_q = nullptr;      // queries
    uint64_t* d_rows = nullptr;   // results
    uint64_t* d_dists = nullptr;
    size_t q_cap = 0, out_cap = 0;
};

static int index_reserve(nps_index* ix, size_t ws, size_t q_words, size_t out_words) {
    if (ws > ix->ws_bytes) {
        if (ix->d_ws) cudaFree(ix->d_ws);
        ix->d_ws = nullptr;
        ix->ws_bytes = 0;
        NPS_CUDA(cudaMalloc(&ix->d_ws, ws));
        ix->ws_bytes = ws;
    }
    if (q_words > ix->q_cap) {
        if (ix->d_q) cudaFree(ix->d_q);
        ix->d_q = nullptr;
        ix->q_cap = 0;
        NPS_CUDA(cudaMalloc(&ix->d_q, q_words * 8));
        ix->q_cap = q_words;
    }
    if (out_words > ix->out_cap) {
        if (ix->d_rows) cudaFree(ix->d_rows);
        if (ix->d_dists) cudaFree(ix->d_dists);
        ix->d_rows = ix->d_dists = nullptr;
        ix->out_cap = 0;
        NPS_CUDA(cudaMalloc(&ix->d_rows, out_words * 8));
        NPS_CUDA(cudaMalloc(&ix->d_dists, out_words * 8));
        ix->out_cap = out_words;
    }
    return NPS_OK;
}

int nps_index_create(const uint64_t* host_hashes, uint64_t num_hashes, uint32_t words, int device,
                     nps_index_t** out_index) {
    if (!out_index) return fail(NPS_EINVAL, "out_index is null");
    *out_index = nullptr;
    if (words == 0 || words > NPS_MAX_WORDS) return fail(NPS_EINVAL, "words=%u outside [1, %u]", words, NPS_MAX_WORDS);
    if (!host_hashes && num_hashes) return fail(NPS_EINVAL, "host_hashes is null");
    DeviceGuard guard(device);
    nps_index* ix = new (std::nothrow) nps_index();
    if (!ix) return fail(NPS_EINVAL, "out of host memory");
    ix->device = device;
    ix->num_hashes = num_hashes;
    ix->words = words;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    const size_t bytes = (size_t)num_hashes * words * 8;
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_hashes, bytes ? bytes : 16);
    if (e == cudaSuccess && bytes)
        e = cudaMemcpyAsync(ix->d_hashes, host_hashes, bytes, cudaMemcpyHostToDevice, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) {
        nps_index_destroy(ix);
        return fail(NPS_ECUDA, "nps_index_create: %s", cudaGetErrorString(e));
    }
    *out_index = ix;
    return NPS_OK;
}

void nps_index_destroy(nps_index_t* ix) {
    if (!ix) return;
    DeviceGuard guard(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    if (ix->d_hashes) cudaFree(ix->d_hashes);
    if (ix->d_ws) cudaFree(ix->d_ws);
    if (ix->d_q) cudaFree(ix->d_q);
    if (ix->d_rows) cudaFree(ix->d_rows);
    if (ix->d_dists) cudaFree(ix->d_dists);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
}

int nps_index_query_top_n(nps_index_t* ix, const uint64_t* host_queries, uint32_t num_queries, uint32_t k, int order,
                          uint64_t* host_rows, uint64_t* host_dists) {
    if (!ix) return fail(NPS_EINVAL, "index is null");
    if (num_queries == 0) return NPS_OK;
    if (!host_queries || !host_rows) return fail(NPS_EINVAL, "null pointer");
    if (k == 0 || k > NPS_MAX_K) return fail(NPS_EINVAL, "k=%u outside [1, %u]", k, NPS_MAX_K);
    if (order != NPS_ORDER_CANONICAL && order != NPS_ORDER_REFERENCE) return fail(NPS_EINVAL, "bad order %d", order);
    DeviceGuard guard(ix->device);
    // the reference-order path materialises [batch, N] distances: bound the batch
    uint32_t batch = num_queries;
    if (order == NPS_ORDER_REFERENCE) {
        const size_t per_q = nps_hamming_top_n_reference_workspace_bytes(ix->num_hashes, ix->words, 1, k);
        const size_t budget = (size_t)1 << 30;
        const size_t fit = per_q ? budget / per_q : num_queries;
        batch = (uint32_t)(fit < 1 ? 1 : (fit < num_queries ? fit : num_queries));
    }
    for (uint32_t q0 = 0; q0 < num_queries; q0 += batch) {
        const uint32_t nq = (num_queries - q0) < batch ? (num_queries - q0) : batch;
        const size_t ws = order == NPS_ORDER_CANONICAL
                              ? nps_hamming_top_n_workspace_bytes(ix->num_hashes, ix->words, nq, k)
                              : nps_hamming_top_n_reference_workspace_bytes(ix->num_hashes, ix->words, nq, k);
        if (order == NPS_ORDER_CANONICAL && ws == 0) return NPS_EINVAL;   // message set by the planner
        int rc = index_reserve(ix, ws, (size_t)nq * ix->words, (size_t)nq * k);
        if (rc) return rc;
        NPS_CUDA(cudaMemcpyAsync(ix->d_q, host_queries + (size_t)q0 * ix->words, (size_t)nq * ix->words * 8,
                                 cudaMemcpyHostToDevice, ix->stream));
        if (order == NPS_ORDER_CANONICAL)
            rc = nps_hamming_top_n(ix->d_hashes, ix->num_hashes, ix->words, ix->d_q, nq, k, 0, ix->d_rows, ix->d_dists,
                                   ix->d_ws, ix->ws_bytes, ix->stream);
        else
            rc = nps_hamming_top_n_reference(ix->d_hashes, ix->num_hashes, ix->words, ix->d_q, nq, k, ix->d_rows,
                                             ix->d_dists, ix->d_ws, ix->ws_bytes, ix->stream);
        if (rc) return rc;
        NPS_CUDA(cudaMemcpyAsync(host_rows + (size_t)q0 * k, ix->d_rows, (size_t)nq * k * 8, cudaMemcpyDeviceToHost,
                                 ix->stream));
        if (host_dists)
            NPS_CUDA(cudaMemcpyAsync(host_dists + (size_t)q0 * k, ix->d_dists, (size_t)nq * k * 8,
                                     cudaMemcpyDeviceToHost, ix->stream));
        NPS_CUDA(cudaStreamSynchronize(ix->stream));
    }
    return NPS_OK;
}

int nps_index_hamming(nps_index_t* ix, const uint64_t* host_query, uint64_t* host_out) {
    if (!ix) return fail(NPS_EINVAL, "index is null");
    if (!host_query || (!host_out && ix->num_hashes)) return fail(NPS_EINVAL, "null pointer");
    if (ix->num_hashes == 0) return NPS_OK;
    DeviceGuard guard(ix->device);
    int rc = index_reserve(ix, (size_t)ix->num_hashes * 8, ix->words, 0);
    if (rc) return rc;
    NPS_CUDA(cudaMemcpyAsync(ix->d_q, host_query, (size_t)ix->words * 8, cudaMemcpyHostToDevice, ix->stream));
    rc = nps_hamming(ix->d_hashes, ix->num_hashes, ix->words, ix->d_q, 1, ix->d_ws, 8, ix->stream);
    if (rc) return rc;
    NPS_CUDA(cudaMemcpyAsync(host_out, ix->d_ws, (size_t)ix->num_hashes * 8, cudaMemcpyDeviceToHost, ix->stream));
    NPS_CUDA(cudaStreamSynchronize(ix->stream));
    return NPS_OK;
}

int nps_host_hamming_top_k(const uint64_t* hashes, const uint64_t* query, uint64_t num_hashes, uint64_t query_len,
                           uint32_t k, uint64_t* best_rows) {
    if (query_len == 0 || query_len > NPS_MAX_WORDS) return fail(NPS_EINVAL, "query_len outside [1, %u]", NPS_MAX_WORDS);
    nps_index_t* ix = nullptr;
    int rc = nps_index_create(hashes, num_hashes, (uint32_t)query_len, 0, &ix);
    if (rc) return rc;
    rc = nps_index_query_top_n(ix, query, 1, k, NPS_ORDER_REFERENCE, best_rows, nullptr);
    nps_index_destroy(ix);
    return rc;
}

int nps_host_hamming(const uint64_t* hashes, const uint64_t* query, uint64_t num_hashes, uint64_t query_len,
                     uint64_t* out) {
    if (query_len == 0 || query_len > NPS_MAX_WORDS) return fail(NPS_EINVAL, "query_len outside [1, %u]", NPS_MAX_WORDS);
    nps_index_t* ix = nullptr;
    int rc = nps_index_create(hashes, num_hashes, (uint32_t)query_len, 0, &ix);
    if (rc) return rc;
    rc = nps_index_hamming(ix, query, out);
    nps_index_destroy(ix);
    return rc;
}

}  // extern "C"
