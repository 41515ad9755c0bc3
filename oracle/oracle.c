/*
 * oracle.c — CPU restatement of np-sims' LSH search path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the B200 kernels.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.  Nothing under np-sims_b200/ imports, links or calls it: the
 * product path is CUDA-only and fails loudly without its extension.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * against (a) the 24 known-answer cases of reference test/test_hamming_top_n.py,
 * (b) the distance/popcount cases of reference test/test_hamming.py, and
 * (c) golden vectors produced by the reference itself (oracle/_ref, built from
 * /root/reference/np_sims/ufunc.c; generator tests/golden/make_golden.py).
 *
 * Plain C, no Python/NumPy headers: loaded through ctypes by oracle/oracle.py.
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference).  It is a restatement, not a copy: one generic-k routine
 * replaces the reference's macro-expanded k=10/k=1000 x W=1..10 bodies.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NONE UINT64_MAX

/* np_sims/ufunc.c:31-37 — population count of one 64-bit word.
 * Written as the SWAR ladder of np_sims/hamming.py:21-38 so that the oracle does
 * not depend on the compiler builtin it is used to check. */
uint64_t orc_popcount64(uint64_t v)
{
    v -= (v >> 1) & 0x5555555555555555ULL;
    v = (v & 0x3333333333333333ULL) + ((v >> 2) & 0x3333333333333333ULL);
    v = (v + (v >> 4)) & 0x0F0F0F0F0F0F0F0FULL;
    return (v * 0x0101010101010101ULL) >> 56;
}

static inline uint64_t row_distance(const uint64_t *row, const uint64_t *query, uint64_t words)
{
    uint64_t d = 0;
    for (uint64_t w = 0; w < words; ++w)
        d += (uint64_t)__builtin_popcountll(row[w] ^ query[w]);
    return d;
}

/* np_sims/ufunc.c:74-106 (gufunc "(m,n),(n)->(m)") and np_sims/hamming.py:41-61:
 * out[i] = sum_w popcount(hashes[i,w] ^ query[w]); hashes is C-contiguous. */
void orc_hamming(const uint64_t *hashes, const uint64_t *query,
                 uint64_t num_hashes, uint64_t words, uint64_t *out)
{
    for (uint64_t i = 0; i < num_hashes; ++i)
        out[i] = row_distance(hashes + i * words, query, words);
}

/* Same distances through the SWAR popcount (np_sims/hamming.py:21-61), used to
 * cross-check the builtin-popcount path above. */
void orc_hamming_swar(const uint64_t *hashes, const uint64_t *query,
                      uint64_t num_hashes, uint64_t words, uint64_t *out)
{
    for (uint64_t i = 0; i < num_hashes; ++i) {
        uint64_t d = 0;
        for (uint64_t w = 0; w < words; ++w)
            d += orc_popcount64(hashes[i * words + w] ^ query[w]);
        out[i] = d;
    }
}

/* np_sims/ufunc.c:41-66 — elementwise popcount(a ^ b) -> uint8.  The reference
 * accumulates into an uninitialised output (":58  +="); the intended value is
 * the plain popcount, which is what this writes. */
void orc_num_unshared_bits(const uint64_t *a, const uint64_t *b, uint64_t n, uint8_t *out)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = (uint8_t)__builtin_popcountll(a[i] ^ b[i]);
}

/* The reference's fixed-capacity selection queue, for any capacity k.
 *   structure + initial state    np_sims/ufunc.c:108-139
 *   insert / evict rule          np_sims/ufunc.c:150-171 (k=10), :174-195 (k=1000)
 *   scan that feeds it           np_sims/ufunc.c:207-216 (unrolled), :411-441 (rolled)
 *
 * best_rows[k]   (out) row ids in SLOT order, ORC_NONE where a slot was never filled
 * best_scores[k] (out, may be NULL) the distances held in those slots
 *
 * Semantics that matter for parity: a row enters only if its
 * distance is STRICTLY below the current worst; while the queue is not full it is
 * appended; afterwards it overwrites the slot `worst_slot`, where worst_slot is the
 * FIRST slot holding the maximum (strict '>' in the rescan). */
void orc_hamming_top_k_queue(const uint64_t *hashes, const uint64_t *query,
                             uint64_t num_hashes, uint64_t words, uint64_t k,
                             uint64_t *best_rows, uint64_t *best_scores)
{
    uint64_t *scores = (uint64_t *)malloc((k ? k : 1) * sizeof(uint64_t));
    uint64_t filled = 0, worst = ORC_NONE, worst_slot = 0;
    for (uint64_t s = 0; s < k; ++s) { scores[s] = ORC_NONE; best_rows[s] = ORC_NONE; }

    for (uint64_t i = 0; i < num_hashes && k > 0; ++i) {
        const uint64_t d = row_distance(hashes + i * words, query, words);
        if (!(d < worst)) continue;
        const uint64_t slot = (filled < k) ? filled++ : worst_slot;
        best_rows[slot] = i;
        scores[slot] = d;
        worst = 0;
        for (uint64_t s = 0; s < k; ++s)
            if (scores[s] > worst) { worst = scores[s]; worst_slot = s; }
    }
    if (best_scores) memcpy(best_scores, scores, k * sizeof(uint64_t));
    free(scores);
}

/* Canonical contract of the new build (BASELINE.json north_star: "ties broken by
 * lowest index"): the k rows with the smallest (distance, row) pairs, ascending.
 * Equivalent to np.argsort(orc_hamming(...), kind='stable')[:k], which is how
 * np_sims/lsh.py:54 selects (its default sort is merely not stable).
 * Unfilled tail: rows = ORC_NONE, dists = ORC_NONE. */
void orc_hamming_top_k_canonical(const uint64_t *hashes, const uint64_t *query,
                                 uint64_t num_hashes, uint64_t words, uint64_t k,
                                 uint64_t *rows, uint64_t *dists)
{
    uint64_t held = 0;
    for (uint64_t s = 0; s < k; ++s) { rows[s] = ORC_NONE; dists[s] = ORC_NONE; }
    for (uint64_t i = 0; i < num_hashes && k > 0; ++i) {
        const uint64_t d = row_distance(hashes + i * words, query, words);
        if (held == k && d >= dists[k - 1]) continue;   /* later row never beats an equal one */
        uint64_t pos = (held < k) ? held++ : k - 1;
        while (pos > 0 && dists[pos - 1] > d) {
            dists[pos] = dists[pos - 1];
            rows[pos] = rows[pos - 1];
            --pos;
        }
        dists[pos] = d;
        rows[pos] = i;
    }
}

/* np_sims/lsh.py:31-33 (index_one) and :36-44 (index): for each vector, bit j of
 * the signature is (dot(projections[j], x) >= 0) in float64; bit j lives in word
 * j/64 at position j%64 (np.packbits(..., bitorder='little').view(uint64)).
 * NaN compares false -> bit 0; an exact zero -> bit 1.
 * vectors may be float32 (is_f32 != 0) or float64; NumPy upcasts f32 to f64. */
void orc_index(const double *projections, uint64_t num_projections, uint64_t dims,
               const void *vectors, int is_f32, uint64_t num_vectors, uint64_t *hashes)
{
    const uint64_t words = num_projections / 64;
    for (uint64_t v = 0; v < num_vectors; ++v) {
        for (uint64_t w = 0; w < words; ++w) hashes[v * words + w] = 0;
        for (uint64_t j = 0; j < words * 64; ++j) {
            const double *p = projections + j * dims;
            double acc = 0.0;
            if (is_f32) {
                const float *x = (const float *)vectors + v * dims;
                for (uint64_t c = 0; c < dims; ++c) acc += p[c] * (double)x[c];
            } else {
                const double *x = (const double *)vectors + v * dims;
                for (uint64_t c = 0; c < dims; ++c) acc += p[c] * x[c];
            }
            if (acc >= 0.0) hashes[v * words + j / 64] |= (uint64_t)1 << (j % 64);
        }
    }
}

/* Same as orc_index but also returns the float64 dot products (for the
 * |x.w| < tol hash-bit tolerance check of the GPU projection kernel). */
void orc_project(const double *projections, uint64_t num_projections, uint64_t dims,
                 const void *vectors, int is_f32, uint64_t num_vectors, double *dots)
{
    for (uint64_t v = 0; v < num_vectors; ++v)
        for (uint64_t j = 0; j < num_projections; ++j) {
            const double *p = projections + j * dims;
            double acc = 0.0;
            if (is_f32) {
                const float *x = (const float *)vectors + v * dims;
                for (uint64_t c = 0; c < dims; ++c) acc += p[c] * (double)x[c];
            } else {
                const double *x = (const double *)vectors + v * dims;
                for (uint64_t c = 0; c < dims; ++c) acc += p[c] * x[c];
            }
            dots[v * num_projections + j] = acc;
        }
}

/* np_sims/lsh.py:84-91 — two-phase query on signatures: top-`cand` rows by the
 * first front_words words (reference queue, k=cand), gather, re-rank by the full
 * signature with the reference queue (k=k), map back through the candidate ids.
 * Padding ids (ORC_NONE) index the LAST row, as NumPy's wrap-around of
 * uint64(-1) -> -1 does in the reference. */
void orc_two_phase(const uint64_t *hashes, const uint64_t *query,
                   uint64_t num_hashes, uint64_t words, uint64_t front_words,
                   uint64_t cand, uint64_t k, uint64_t *out_rows)
{
    uint64_t *front = (uint64_t *)malloc((num_hashes * front_words + 1) * sizeof(uint64_t));
    uint64_t *cands = (uint64_t *)malloc((cand + 1) * sizeof(uint64_t));
    uint64_t *gathered = (uint64_t *)malloc((cand * words + 1) * sizeof(uint64_t));
    uint64_t *best = (uint64_t *)malloc((k + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < num_hashes; ++i)
        memcpy(front + i * front_words, hashes + i * words, front_words * sizeof(uint64_t));
    orc_hamming_top_k_queue(front, query, num_hashes, front_words, cand, cands, NULL);
    for (uint64_t c = 0; c < cand; ++c) {
        const uint64_t row = (cands[c] == ORC_NONE) ? num_hashes - 1 : cands[c];
        memcpy(gathered + c * words, hashes + row * words, words * sizeof(uint64_t));
    }
    orc_hamming_top_k_queue(gathered, query, cand, words, k, best, NULL);
    for (uint64_t s = 0; s < k; ++s)
        out_rows[s] = (best[s] == ORC_NONE) ? cands[cand - 1] : cands[best[s]];
    free(front); free(cands); free(gathered); free(best);
}
