// searcher.cu — host-level single-query path: the reference's calling pattern, one query per call from
// many host threads over a shared read-only table (benchmark/glove.py:183-199; the GIL is released in
// the ufunc, np_sims/ufunc.c:456-457), made cheap on a GPU.
//
//   nps_searcher_top_k         = gufuncs hamming_top_10 / hamming_top_cand (ufunc.c:444-551), HOST query
//                                signature in, HOST ids out, reference queue order
//   nps_searcher_query_vector  = lsh.query (lsh.py:70-74): HOST vector in -> float64 projection + sign +
//                                pack (lsh.py:31-33) -> the same search, HOST ids out
//
// Every calling thread takes a SLOT: its own stream, pinned staging buffers, device buffers and
// workspace.  The whole call — H2D copy, hashing kernel, K2R prefix + scan/replay kernels
// (ref_scan.cuh), D2H copy of ids and of the overflow flag — is ONE CUDA graph per (slot, shape),
// captured on first use: a call costs one cudaGraphLaunch + one cudaStreamSynchronize, and calls of
// different threads overlap on the device.  Nothing here is process-global mutable state.
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/npsims_b200.h"
#include "common.cuh"
#include "hamming.cuh"
#include "ref_scan.cuh"
#include "replay.cuh"

namespace nps {

// lsh.index_one (lsh.py:31-33) for ONE vector: bit j = (dot(projections[j], x) >= 0) in float64.
// One CTA per output word: 32 warps x 2 projections, lanes stride over the dimensions.
constexpr int kProjOneThreads = 1024;
__global__ void __launch_bounds__(kProjOneThreads) project_one_kernel(const void* __restrict__ x, int x_is_f32, uint32_t dims,
                                                                     const double* __restrict__ Wp,
                                                                     uint64_t* __restrict__ out) {
    extern __shared__ double xs[];
    __shared__ unsigned long long word;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < dims; i += kProjOneThreads)
        xs[i] = x_is_f32 ? (double)reinterpret_cast<const float*>(x)[i] : reinterpret_cast<const double*>(x)[i];
    if (tid == 0) word = 0ull;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const uint32_t bit = (uint32_t)warp * 2u + j;
        const double* w = Wp + ((size_t)blockIdx.x * 64u + bit) * dims;
        double acc = 0.0;
        for (uint32_t i = lane; i < dims; i += 32) acc = fma(__ldg(w + i), xs[i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && acc >= 0.0) atomicOr(&word, 1ull << bit);   // NaN -> 0, exact zero -> 1 (lsh.py:32)
    }
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = word;
}

struct GraphKey {
    int mode;                 // 0: signature query, 1: float32 vector, 2: float64 vector
    uint32_t k, dims, nproj;
    const double* proj;
    bool operator==(const GraphKey& o) const {
        return mode == o.mode && k == o.k && dims == o.dims && nproj == o.nproj && proj == o.proj;
    }
};
struct SlotGraph {
    GraphKey key;
    cudaGraphExec_t exec;
};
struct Slot {
    cudaStream_t stream = nullptr;
    uint8_t* h_in = nullptr;       // pinned
    uint64_t* h_out = nullptr;     // pinned: k ids + 1 word (overflow flag)
    size_t h_in_bytes = 0, h_out_words = 0;
    uint8_t* d_in = nullptr;       // query vector
    uint64_t* d_sig = nullptr;
    uint64_t* d_rows = nullptr;
    void* d_ws = nullptr;
    size_t d_in_bytes = 0, ws_bytes = 0, rows_words = 0;
    void* d_dense = nullptr;       // fallback: dense distances
    size_t dense_bytes = 0;
    std::vector<SlotGraph> graphs;
};

}  // namespace nps

using namespace nps;

struct nps_searcher {
    int device = 0;
    int sm_count = 148, smem_optin = 232448;
    const uint64_t* d_hashes = nullptr;
    uint64_t N = 0;
    uint32_t W = 0;
    std::mutex mu;
    std::vector<Slot*> free_slots, all_slots;
    unsigned long long calls = 0, fallbacks = 0;
};

#define S_CUDA(expr)                                                                                    \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) return nps::fail(NPS_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__));  \
    } while (0)
static int s_fail(int code, const char* msg) { return nps::fail(code, "%s", msg); }

namespace {

struct DeviceGuard {   // the caller's current device is left as it was
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

void slot_free(Slot* s) {
    if (!s) return;
    for (auto& g : s->graphs) cudaGraphExecDestroy(g.exec);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->d_in) cudaFree(s->d_in);
    if (s->d_sig) cudaFree(s->d_sig);
    if (s->d_rows) cudaFree(s->d_rows);
    if (s->d_ws) cudaFree(s->d_ws);
    if (s->d_dense) cudaFree(s->d_dense);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int slot_reserve(nps_searcher* S, Slot* s, size_t in_bytes, uint32_t k, size_t ws_bytes) {
    bool changed = false;
    if (!s->stream) S_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    if (in_bytes > s->h_in_bytes) {
        if (s->h_in) cudaFreeHost(s->h_in);
        if (s->d_in) cudaFree(s->d_in);
        s->h_in = nullptr;
        s->d_in = nullptr;
        s->h_in_bytes = 0;
        const size_t b = in_bytes < 4096 ? 4096 : in_bytes;
        S_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_in), b));
        S_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_in), b));
        s->h_in_bytes = b;
        changed = true;
    }
    if (!s->d_sig) S_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_sig), (size_t)NPS_MAX_WORDS * 8));
    if ((size_t)k + 1 > s->h_out_words) {
        if (s->h_out) cudaFreeHost(s->h_out);
        if (s->d_rows) cudaFree(s->d_rows);
        s->h_out = nullptr;
        s->d_rows = nullptr;
        s->h_out_words = 0;
        const size_t wds = (size_t)(k < 16 ? 16 : k) + 1;
        S_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_out), wds * 8));
        S_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_rows), wds * 8));
        s->h_out_words = wds;
        changed = true;
    }
    if (ws_bytes > s->ws_bytes) {
        if (s->d_ws) cudaFree(s->d_ws);
        s->d_ws = nullptr;
        s->ws_bytes = 0;
        S_CUDA(cudaMalloc(&s->d_ws, ws_bytes));
        S_CUDA(cudaMemsetAsync(s->d_ws, 0, ref_zero_bytes(1), s->stream));   // once: every call leaves it clean
        s->ws_bytes = ws_bytes;
        changed = true;
    }
    if (changed) {   // captured graphs hold the old pointers
        for (auto& g : s->graphs) cudaGraphExecDestroy(g.exec);
        s->graphs.clear();
    }
    return NPS_OK;
}

// enqueue one query on the slot's stream (directly, or under stream capture)
int enqueue(nps_searcher* S, Slot* s, const GraphKey& key, const RefPlan& pl, size_t in_bytes) {
    const uint64_t* d_q = s->d_sig;
    if (key.mode == 0) {
        S_CUDA(cudaMemcpyAsync(s->d_sig, s->h_in, in_bytes, cudaMemcpyHostToDevice, s->stream));
    } else {
        S_CUDA(cudaMemcpyAsync(s->d_in, s->h_in, in_bytes, cudaMemcpyHostToDevice, s->stream));
        project_one_kernel<<<key.nproj / 64, kProjOneThreads, key.dims * sizeof(double), s->stream>>>(
            s->d_in, key.mode == 1, key.dims, key.proj, s->d_sig);
        count_launch();
        S_CUDA(cudaGetLastError());
    }
    S_CUDA(run_ref_top_n(pl, S->d_hashes, S->N, S->W, d_q, 1, key.k, s->d_rows, nullptr, s->d_ws, /*zero_first=*/false, s->stream));
    S_CUDA(cudaMemcpyAsync(s->h_out, s->d_rows, (size_t)key.k * 8, cudaMemcpyDeviceToHost, s->stream));
    S_CUDA(cudaMemcpyAsync(s->h_out + key.k, s->d_ws, 4, cudaMemcpyDeviceToHost, s->stream));
    return NPS_OK;
}

int run_query(nps_searcher* S, const GraphKey& key, const void* host_in, size_t in_bytes, uint64_t* host_rows) {
    if (!S) return s_fail(NPS_EINVAL, "searcher is null");
    if (!host_in || !host_rows) return s_fail(NPS_EINVAL, "null pointer");
    if (key.k == 0 || key.k > NPS_MAX_K) return s_fail(NPS_EINVAL, "k outside [1, 2048]");
    const RefPlan pl = ref_plan(S->N, S->W, key.k, S->sm_count, S->smem_optin);
    if (!pl.applicable) return s_fail(NPS_EINVAL, "shape not supported by the single-query path (words > 63?)");
    DeviceGuard guard(S->device);
    Slot* s = nullptr;
    {
        std::lock_guard<std::mutex> lock(S->mu);
        if (!S->free_slots.empty()) {
            s = S->free_slots.back();
            S->free_slots.pop_back();
        }
        ++S->calls;
    }
    if (!s) {
        s = new (std::nothrow) Slot();
        if (!s) return s_fail(NPS_EINVAL, "out of host memory");
        std::lock_guard<std::mutex> lock(S->mu);
        S->all_slots.push_back(s);
    }
    struct Release {
        nps_searcher* S;
        Slot* s;
        ~Release() {
            std::lock_guard<std::mutex> lock(S->mu);
            S->free_slots.push_back(s);
        }
    } release{S, s};
    int rc = slot_reserve(S, s, in_bytes, key.k, ref_zero_bytes(1) + pl.ws_stride);
    if (rc) return rc;
    memcpy(s->h_in, host_in, in_bytes);
    cudaGraphExec_t exec = nullptr;
    for (auto& g : s->graphs)
        if (g.key == key) exec = g.exec;
    if (!exec) {
        // first call of this shape on this slot: run it directly (also sets the kernels' shared-memory
        // attributes), then capture the same sequence for every later call
        rc = enqueue(S, s, key, pl, in_bytes);
        if (rc) return rc;
        S_CUDA(cudaStreamSynchronize(s->stream));
        cudaGraph_t graph = nullptr;
        S_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueue(S, s, key, pl, in_bytes);
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        S_CUDA(e);
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        S_CUDA(e);
        s->graphs.push_back(SlotGraph{key, exec});
    } else {
        S_CUDA(cudaGraphLaunch(exec, s->stream));
        count_launch((key.mode ? 1u : 0u) + 1u + (pl.num_chunks ? 1u : 0u));   // kernels inside the graph
        S_CUDA(cudaStreamSynchronize(s->stream));
    }
    if ((uint32_t)s->h_out[key.k] != 0u) {
        // a chunk's segment overflowed (adversarial row order): K5 + dense replay, exact
        const int elt = 2;
        const size_t need = ((size_t)S->N * elt + 255) & ~(size_t)255;
        if (need > s->dense_bytes) {
            if (s->d_dense) cudaFree(s->d_dense);
            s->d_dense = nullptr;
            s->dense_bytes = 0;
            S_CUDA(cudaMalloc(&s->d_dense, need));
            s->dense_bytes = need;
        }
        S_CUDA(launch_hamming_dist(S->d_hashes, S->N, S->W, s->d_sig, 1, s->d_dense, elt, S->sm_count, s->stream));
        S_CUDA(launch_replay_queue(s->d_dense, elt, S->N, 1, key.k, s->d_rows, nullptr, s->stream));
        S_CUDA(cudaMemcpyAsync(s->h_out, s->d_rows, (size_t)key.k * 8, cudaMemcpyDeviceToHost, s->stream));
        S_CUDA(cudaStreamSynchronize(s->stream));
        std::lock_guard<std::mutex> lock(S->mu);
        ++S->fallbacks;
    }
    memcpy(host_rows, s->h_out, (size_t)key.k * 8);
    return NPS_OK;
}

}  // namespace

extern "C" {

int nps_searcher_create(const uint64_t* d_hashes, uint64_t num_hashes, uint32_t words, int device,
                        nps_searcher_t** out) {
    if (!out) return s_fail(NPS_EINVAL, "out is null");
    *out = nullptr;
    if (words == 0 || words > kRefMaxWords) return s_fail(NPS_EINVAL, "words outside [1, 63]");
    if (!d_hashes && num_hashes) return s_fail(NPS_EINVAL, "d_hashes is null");
    if (reinterpret_cast<uintptr_t>(d_hashes) & 15u) return s_fail(NPS_EINVAL, "d_hashes must be 16-byte aligned");
    DeviceGuard guard(device);
    nps_searcher* S = new (std::nothrow) nps_searcher();
    if (!S) return s_fail(NPS_EINVAL, "out of host memory");
    S->device = device;
    S->d_hashes = d_hashes;
    S->N = num_hashes;
    S->W = words;
    if (cudaDeviceGetAttribute(&S->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaDeviceGetAttribute(&S->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) {
        delete S;
        return s_fail(NPS_ECUDA, "cudaDeviceGetAttribute failed (no CUDA device?)");
    }
    *out = S;
    return NPS_OK;
}

void nps_searcher_destroy(nps_searcher_t* S) {
    if (!S) return;
    DeviceGuard guard(S->device);
    for (Slot* s : S->all_slots) slot_free(s);
    delete S;
}

int nps_searcher_top_k(nps_searcher_t* S, const uint64_t* host_query, uint32_t k, uint64_t* host_rows) {
    if (!S) return s_fail(NPS_EINVAL, "searcher is null");
    GraphKey key{0, k, 0, 0, nullptr};
    return run_query(S, key, host_query, (size_t)S->W * 8, host_rows);
}

int nps_searcher_query_vector(nps_searcher_t* S, const void* host_vector, int vector_is_f32, uint32_t dims,
                              const double* d_projections, uint32_t num_projections, uint32_t k, uint64_t* host_rows) {
    if (!S) return s_fail(NPS_EINVAL, "searcher is null");
    if (!d_projections || dims == 0 || dims > 6000) return s_fail(NPS_EINVAL, "bad projections / dims (1..6000)");
    if (num_projections != S->W * 64u) return s_fail(NPS_EINVAL, "num_projections must equal 64 * words of the table");
    GraphKey key{vector_is_f32 ? 1 : 2, k, dims, num_projections, d_projections};
    return run_query(S, key, host_vector, (size_t)dims * (vector_is_f32 ? 4 : 8), host_rows);
}

int nps_searcher_stats(nps_searcher_t* S, unsigned long long* calls, unsigned long long* fallbacks, uint32_t* slots) {
    if (!S) return s_fail(NPS_EINVAL, "searcher is null");
    std::lock_guard<std::mutex> lock(S->mu);
    if (calls) *calls = S->calls;
    if (fallbacks) *fallbacks = S->fallbacks;
    if (slots) *slots = (uint32_t)S->all_slots.size();
    return NPS_OK;
}

}  // extern "C"
