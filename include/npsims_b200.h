/*
 * npsims_b200.h — C ABI of libnpsims_b200.so: the B200 (sm_100a) replacement for the native
 * hot path of softwaredoug/np-sims (np_sims/ufunc.c + the hashing loop of np_sims/lsh.py).
 *
 * Plain C: raw pointers, sizes and an opaque stream handle (a cudaStream_t passed as
 * void*).  No torch, NumPy or Python types.  Every function returns 0 on success and a
 * negative NPS_E* code on failure; nps_last_error() returns a thread-local message.
 * The library never falls back to the CPU: without a CUDA device every compute entry
 * point fails with NPS_ECUDA.
 *
 * Two levels:
 *   A. nps_host_*  / nps_index_*  take HOST buffers, like the reference's gufunc inner loops,
 *      and are what a maintainer binds from ufunc.c / ctypes;
 *   B. nps_*  (device level) take DEVICE pointers and a stream; the Python host layer
 *      (np-sims_b200/np_sims) calls these with torch tensors as the buffer carrier.
 *
 * Citations "ufunc.c:LINE" / "lsh.py:LINE" are relative to /root/reference/np_sims/.
 */
#ifndef NPSIMS_B200_H
#define NPSIMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

#define NPS_OK 0
#define NPS_EINVAL (-1)    /* bad argument (shape, k, alignment, null pointer) */
#define NPS_ECUDA (-2)     /* CUDA runtime error / no device */
#define NPS_EWORKSPACE (-3) /* workspace too small */

#define NPS_MAX_K 2048u        /* largest supported top-n */
#define NPS_MAX_WORDS 1024u    /* largest signature width in 64-bit words (65536 bits) */
#define NPS_NONE UINT64_MAX    /* padding id when fewer than k rows exist (ufunc.c:129-132) */

/* top-n result order */
#define NPS_ORDER_CANONICAL 0  /* ascending (distance, row id): "ties broken by lowest index" */
#define NPS_ORDER_REFERENCE 1  /* the reference queue's slot order and tie survivors, bit-exact
                                  with ufunc.c:150-195 */

int nps_version(void);
const char *nps_last_error(void);
/* number of CUDA devices visible, or NPS_ECUDA */
int nps_device_count(void);

/* Measurement aids (bench.py): with profiling on, every nps_hamming_top_n* call brackets its
 * scan (K2) and merge (K3) launches with CUDA events on the caller's stream;
 * nps_profile_last() waits for them and returns the two device durations in milliseconds.
 * nps_last_scan_plan() reports how the last call was tiled.  Off by default.
 * Threading: the profiling switch, the forced kernel families (nps_set_scan_path,
 * nps_set_hash_path) and every nps_last_* diagnostic are PER HOST THREAD (thread_local): they
 * describe and affect only calls made by the same thread.  All compute entry points are
 * re-entrant across host threads as long as each call uses its own stream / workspace / outputs. */
typedef struct nps_scan_plan {
    uint32_t tile_rows, tile_bytes, stages, q_tile, num_qtiles, chunk_rows, num_chunks, cap,
        smem_bytes, grid, specialised, rows_per_thread;
} nps_scan_plan_t;
int nps_profile_enable(int on);
int nps_profile_last(float *scan_ms, float *merge_ms);
int nps_last_scan_plan(nps_scan_plan_t *out);
/* Kernel family of the top-n scan.  AUTO: batches of >= 64 queries over >= 4096 rows, and
 * batches of >= 16 queries once rows * queries * words >= 1e9, run on the tensor cores
 * (tcgen05 kind::mxf4 block-scaled GEMM over E2M1 +-1 operands with unit scales, FP32
 * accumulation in TMEM: exact integers, csrc/mma_scan.cu); everything else on the XOR+POPC
 * streaming scan (csrc/scan.cuh).  Forcing a family is for tests and measurements; results are
 * identical. */
#define NPS_SCAN_AUTO 0
#define NPS_SCAN_POPC 1
#define NPS_SCAN_MMA 2
int nps_set_scan_path(int path);
int nps_last_scan_path(void);
/* Tensor-core scan: CTA pairs (tcgen05 cta_group::2: one M = 256 MMA per two SMs of a TPC, each CTA
 * holding half of the query tile).  mode 0: never; 1 (default): where the single-CTA plan is squeezed
 * by shared memory (signatures from ~1024 bits); 2: always.  Process-wide; for measurements and
 * tests — results are identical. */
int nps_set_mma_pair(int mode);
typedef struct nps_mma_plan {
    uint32_t n_tile, num_qtiles, cap, growth, a_stages, smem_bytes, num_phases, last_tiles,
        last_tiles_per_chunk, last_num_chunks, last_grid, pair;
} nps_mma_plan_t;
int nps_last_mma_plan(nps_mma_plan_t *out);
/* After a profiled call that took the tensor-core path: device duration of each phase's scan
 * launch and the number of 128-row tiles it covered (phase 0 = dense bootstrap). */
int nps_profile_last_phases(float *scan_ms, uint32_t *tiles, uint32_t max_phases, uint32_t *num_phases);
/* Tensor-core path only: number of queries whose candidate buffer overflowed during the call
 * that used d_workspace (0 in all but adversarial inputs).  Synchronises the stream.  Purely a
 * diagnostic: when the count is non-zero the call has already been repaired ON THE DEVICE — every
 * tensor-core top-n call enqueues the XOR+POPC scan + merge behind the phases, each launch gated
 * on this counter, so results are exact either way and no caller action is needed. */
int nps_top_n_overflow_count(const void *d_workspace, void *stream, uint32_t *out_count);
/* cumulative number of CUDA kernels this library has launched in this process */
unsigned long long nps_launch_count(void);

/* ------------------------------------------------------------------ A. host level ------- */

/* Drop-in for the bodies behind gufunc `hamming_top_10` / `hamming_top_cand`
 * (ufunc.c:411-441 hamming_top_{10,1000}_default and :207-406 unrolled variants):
 * hashes[num_hashes*query_len], query[query_len] and best_rows[k] are HOST pointers,
 * C-contiguous uint64.  Result is bit-identical to the reference (slot order, UINT64_MAX
 * padding).  Uploads the table, runs K5 + the queue replay on device 0, downloads k ids. */
int nps_host_hamming_top_k(const uint64_t *hashes, const uint64_t *query, uint64_t num_hashes,
                           uint64_t query_len, uint32_t k, uint64_t *best_rows);

/* Drop-in for the body of gufunc `hamming` (ufunc.c:74-106): out[num_hashes] distances. */
int nps_host_hamming(const uint64_t *hashes, const uint64_t *query, uint64_t num_hashes,
                     uint64_t query_len, uint64_t *out);

/* A signature table resident in HBM (what lsh.index returns, lsh.py:36-44), so that
 * repeated queries do not re-upload it. */
typedef struct nps_index nps_index_t;
int nps_index_create(const uint64_t *host_hashes, uint64_t num_hashes, uint32_t words, int device,
                     nps_index_t **out_index);
void nps_index_destroy(nps_index_t *index);
/* Batched top-n: host_queries[num_queries*words] -> host_rows[num_queries*k] (and
 * host_dists[num_queries*k] if not NULL).  order = NPS_ORDER_*.  Copies queries H2D, runs
 * K2+K3 (canonical) or K5+replay (reference), copies results D2H; returns when done. */
int nps_index_query_top_n(nps_index_t *index, const uint64_t *host_queries, uint32_t num_queries,
                          uint32_t k, int order, uint64_t *host_rows, uint64_t *host_dists);
/* Full distances of one query against the resident table: host_out[num_hashes]. */
int nps_index_hamming(nps_index_t *index, const uint64_t *host_query, uint64_t *host_out);

/* The reference's own calling pattern — one query per call, many host threads over a shared
 * read-only table (benchmark/glove.py:183-199; the GIL is released at ufunc.c:456-457) — over a table
 * that is ALREADY resident in HBM (d_hashes [num_hashes, words], 16-byte aligned, owned by the caller
 * and alive as long as the searcher).  Thread-safe: every calling thread takes its own slot (stream,
 * pinned staging, workspace); a call is one CUDA-graph launch (H2D copy, [hash kernel,] K2R prefix +
 * scan/replay kernels, D2H copy) and one stream synchronisation.  Results are bit-identical to the
 * reference queue (slot order, NPS_NONE padding).  words <= 63.
 *   nps_searcher_top_k        : gufuncs hamming_top_10 (k = 10) / hamming_top_cand (k = 1000),
 *                               ufunc.c:444-551; host_query[words] -> host_rows[k]
 *   nps_searcher_query_vector : lsh.query (lsh.py:70-74): host_vector[dims] float32 / float64 ->
 *                               float64 projection + sign + pack (lsh.py:31-33, d_projections
 *                               [num_projections, dims] float64 in HBM) -> the same search */
typedef struct nps_searcher nps_searcher_t;
int nps_searcher_create(const uint64_t *d_hashes, uint64_t num_hashes, uint32_t words, int device,
                        nps_searcher_t **out_searcher);
void nps_searcher_destroy(nps_searcher_t *searcher);
int nps_searcher_top_k(nps_searcher_t *searcher, const uint64_t *host_query, uint32_t k, uint64_t *host_rows);
int nps_searcher_query_vector(nps_searcher_t *searcher, const void *host_vector, int vector_is_f32,
                              uint32_t dims, const double *d_projections, uint32_t num_projections,
                              uint32_t k, uint64_t *host_rows);
/* diagnostics: calls served, calls that took the dense fallback, slots created */
int nps_searcher_stats(nps_searcher_t *searcher, unsigned long long *calls, unsigned long long *fallbacks,
                       uint32_t *slots);

/* ------------------------------------------------------------------ B. device level ----- */

/* K5 — gufunc `hamming` (ufunc.c:74-106), batched: d_out[num_queries, num_hashes] of
 * uint16/uint32/uint64 (out_bytes = 2/4/8).  d_hashes [num_hashes, words] row-major. */
int nps_hamming(const uint64_t *d_hashes, uint64_t num_hashes, uint32_t words,
                const uint64_t *d_queries, uint32_t num_queries, void *d_out, int out_bytes,
                void *stream);

/* K2+K3 — gufuncs `hamming_top_10` / `hamming_top_cand` (ufunc.c:444-551) generalised to any
 * k <= NPS_MAX_K and a batch of queries, canonical order.  d_out_rows/d_out_dists
 * [num_queries, k] uint64 (d_out_dists may be NULL); ids are row + row_base; padding NPS_NONE.
 * d_hashes must be 16-byte aligned.  d_workspace: nps_hamming_top_n_workspace_bytes(). */
size_t nps_hamming_top_n_workspace_bytes(uint64_t num_hashes, uint32_t words, uint32_t num_queries,
                                         uint32_t k);
int nps_hamming_top_n(const uint64_t *d_hashes, uint64_t num_hashes, uint32_t words,
                      const uint64_t *d_queries, uint32_t num_queries, uint32_t k,
                      uint64_t row_base, uint64_t *d_out_rows, uint64_t *d_out_dists,
                      void *d_workspace, size_t workspace_bytes, void *stream);

/* Same search, but returns the per-query candidate KEYS (distance << 40 | row id, ascending,
 * NPS_NONE padded) instead of decoded ids: the unit exchanged between GPUs (all-gather) and
 * merged by nps_top_n_merge. d_out_keys [num_queries, k]. */
int nps_hamming_top_n_keys(const uint64_t *d_hashes, uint64_t num_hashes, uint32_t words,
                           const uint64_t *d_queries, uint32_t num_queries, uint32_t k,
                           uint64_t row_base, uint64_t *d_out_keys, void *d_workspace,
                           size_t workspace_bytes, void *stream);

/* K4 — merge `num_lists` candidate-key lists per query (layout [num_queries, num_lists, k] or,
 * with list_major != 0, [num_lists, num_queries, k] as an all-gather leaves them) into the
 * final [num_queries, k] ids/distances.  No reference equivalent (the reference is
 * single-process); equals running the search on the concatenated table.  Preconditions: every
 * list ascending and NPS_NONE padded, keys unique ACROSS the lists of a query (true for shards
 * of one table: the row id is part of the key) — a key present in two lists leaves output
 * slots unwritten. */
size_t nps_top_n_merge_workspace_bytes(uint32_t num_queries, uint32_t num_lists, uint32_t k);
int nps_top_n_merge(const uint64_t *d_keys, uint32_t num_lists, uint32_t num_queries, uint32_t k,
                    int list_major, uint64_t *d_out_rows, uint64_t *d_out_dists, void *d_workspace,
                    size_t workspace_bytes, void *stream);

/* Reference-exact order (NPS_ORDER_REFERENCE) — gufuncs `hamming_top_10` / `hamming_top_cand`
 * (ufunc.c:444-551) with the queue of ufunc.c:108-195 reproduced bit for bit: d_out_rows
 * [num_queries, k] in SLOT order, NPS_NONE where a slot was never filled; d_out_scores (nullable)
 * the slot distances.  K2R (csrc/ref_scan.cuh): a prefix kernel bounds the queue's worst, a
 * streaming scan at HBM speed keeps the rows that can ever enter the queue, the last CTA replays
 * them in row order — two launches (one for tables <= 32768 rows).  Signatures wider than 63 words,
 * unaligned tables and (gated on a device flag, no host round trip) adversarial row orders take
 * K5 + the dense replay (csrc/replay.cu).  d_workspace: 256-byte aligned,
 * nps_hamming_top_n_reference_workspace_bytes(num_hashes, words, num_queries, k) bytes. */
size_t nps_hamming_top_n_reference_workspace_bytes(uint64_t num_hashes, uint32_t words,
                                                   uint32_t num_queries, uint32_t k);
int nps_hamming_top_n_reference(const uint64_t *d_hashes, uint64_t num_hashes, uint32_t words,
                                const uint64_t *d_queries, uint32_t num_queries, uint32_t k,
                                uint64_t *d_out_rows, uint64_t *d_out_scores, void *d_workspace,
                                size_t workspace_bytes, void *stream);

/* gufunc `num_unshared_bits` (ufunc.c:41-66), elementwise popcount(a ^ b) -> uint8 with
 * NumPy-style cyclic broadcasting: d_out[i] = popc(d_a[i % len_a] ^ d_b[i % len_b]), i < n.
 * Writes the popcount (the reference's `+=` into uninitialised memory is a bug, not parity). */
int nps_num_unshared_bits(const uint64_t *d_a, uint64_t len_a, const uint64_t *d_b, uint64_t len_b,
                          uint64_t n, uint8_t *d_out, void *stream);

/* K1a — lsh.index / lsh.index_one (lsh.py:31-44) in the reference's own arithmetic: float64
 * dot products on the FP64 pipe, bit j of row i = (dot(projections[j], vectors[i]) >= 0),
 * stored at word j/64, bit j%64.  d_vectors [num_vectors, dims] float32 (vectors_are_f32 != 0)
 * or float64; d_projections [num_projections, dims] float64; num_projections % 64 == 0.
 * d_out [num_vectors, num_projections/64] (nullable); d_dots [num_vectors, num_projections]
 * float64 (nullable) receives the raw projections for tolerance analysis. */
int nps_project_sign_pack_f64(const void *d_vectors, int vectors_are_f32, uint64_t num_vectors,
                              uint32_t dims, const double *d_projections, uint32_t num_projections,
                              uint64_t *d_out, double *d_dots, void *stream);

/* K6 — second phase of the two-phase search (lsh.py:84-91: `hashes[candidates]` gathered, then
 * hamming_top_10 on the copy, then `candidates[best]`): for every query the k candidates with the
 * smallest FULL-width distance, ascending (distance, row id), reading the candidate rows in place.
 * d_candidates [num_queries, num_candidates] global row ids (e.g. nps_hamming_top_n on the front
 * words); ids of UINT64_MAX or >= num_hashes are ignored; num_candidates <= 4096. */
int nps_rerank_top_n(const uint64_t *d_hashes, uint64_t num_hashes, uint32_t words, const uint64_t *d_queries,
                     uint32_t num_queries, const uint64_t *d_candidates, uint32_t num_candidates, uint32_t k,
                     uint64_t *d_out_rows, uint64_t *d_out_dists, void *stream);

/* K1T — the same hashing for float32 vectors on the tensor cores: TMA-fed tcgen05 kind::tf32
 * GEMM with a sign/bit-pack epilogue.  Both operands are split into two TF32 terms (x = xh + xl in
 * the kernel, w = wh + wl by nps_project_split_projections) and three products are accumulated, so
 * an accumulator is exact to (8 + dims / 4) * 2^-22 * |x_i| * projection_norm_max; the (row,
 * projection) pairs below that (0.03 % at dims = 300) and all rows with a zero / tiny / huge /
 * non-finite norm are recomputed in float64 — by fixer warps inside the hashing kernel (BF16 encoding) or,
 * listed, by a fix-up kernel — so the signatures equal those of nps_project_sign_pack_f64.  Needs dims % 4 == 0 and 16-byte aligned inputs.
 * d_projections_split: nps_project_split_bytes() bytes filled by nps_project_split_projections (do it
 * once per projection matrix); d_projections_f64: the float64 matrix (fix-up);
 * projection_norm_max = max_j |projections[j]| (1.0 for lsh.create_projections).
 * If the fix-up list overflows (degenerate input) the call is redone in float64 by a gated launch
 * on the device — no host round trip.  nps_project_sign_pack_stats (diagnostic; synchronises the
 * stream) returns how many pairs went to the fix-up list and its capacity (listed > capacity: the call was redone
 * in float64); nps_project_sign_pack_stats_in_kernel how many more the fixer warps recomputed inside the kernel. */
/* size in bytes of the split buffer for nps_project_split_projections / nps_project_sign_pack (both
 * operand encodings: float32 [2, B, dims] TF32 parts, then bfloat16 [2, B, dims rounded up to 8], then the
 * projections' prefix norms max_j |w_j[0 : 16 i)|, one float32 per 16 dims, for the BF16 kernel's tolerance) */
size_t nps_project_split_bytes(uint32_t num_projections, uint32_t dims);
/* which operand encoding nps_project_sign_pack uses: BF16 x 2 on kind::f16 (AUTO, the default; tolerance
 * (52 + 0.006 dims) * 2^-22 |x| max|w| + 2.68 * 2^-22 * sum_i |x[0 : 16 i)| max_j |w_j[0 : 16 i)|, at most
 * (52 + 0.17 dims) * 2^-22 |x| max|w|) or TF32 x 2 on kind::tf32 (tolerance (8 + dims/4) * 2^-22 |x| max|w|) */
#define NPS_HASH_AUTO 0
#define NPS_HASH_TF32 1
#define NPS_HASH_BF16 2
int nps_set_hash_path(int path);
int nps_project_split_projections(const double *d_projections_f64, uint32_t num_projections, uint32_t dims,
                                  float *d_out_split, void *stream);
size_t nps_project_sign_pack_workspace_bytes(uint64_t num_vectors, uint32_t num_projections);
int nps_project_sign_pack(const float *d_vectors, uint64_t num_vectors, uint32_t dims,
                          const float *d_projections_split, const double *d_projections_f64,
                          uint32_t num_projections, float projection_norm_max, uint64_t *d_out,
                          void *d_workspace, size_t workspace_bytes, void *stream);
int nps_project_sign_pack_stats(const void *d_workspace, uint64_t num_vectors, uint32_t num_projections,
                                void *stream, uint32_t *recomputed, uint32_t *capacity);
int nps_project_sign_pack_stats_in_kernel(const void *d_workspace, uint64_t num_vectors, void *stream,
                                          uint32_t *in_kernel);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* NPSIMS_B200_H */
