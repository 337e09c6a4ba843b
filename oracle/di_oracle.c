/*
 * di_oracle.c — CPU restatement of the reference's inverted-index hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it. The product path
 * (improving-learned-index_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks every function here against
 * outputs of the reference's own Python code (quantize_file, InvertedIndexCreator,
 * InvertedIndex.score, SparseSearch.search) recorded by oracle/make_golden.py into
 * tests/golden/.
 *
 * Each function cites the reference file:line (relative to /root/reference) it follows.
 * Plain C99 + OpenMP (only for the multi-query scorer used as the timed CPU baseline).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------------
 * Quantization — src/deep_impact/indexing/quantize.py
 * ------------------------------------------------------------------------------- */

/* quantize.py:17-24 find_max_value: running max over every float(score), seeded with 0. */
double dio_find_max(const double *scores, int64_t n)
{
    double m = 0.0;
    for (int64_t i = 0; i < n; ++i)
        if (scores[i] > m) m = scores[i];
    return m;
}

/* quantize.py:37 scale = ((1 << 8) - 1) / max_val   (IMPACT_SCORE_QUANTIZATION_BITS = 8,
 * src/utils/defaults.py:26), evaluated in float64. */
double dio_scale(double max_val) { return 255.0 / max_val; }

/* quantize.py:13-14 quantize(value, scale) = int(value * scale): one float64 multiply,
 * then truncation toward zero. The caller keeps the posting iff the result is > 0
 * (quantize.py:45). Results are NOT clamped to 255 here (the reference does not clamp). */
void dio_quantize(const double *scores, int64_t n, double scale, int64_t *out)
{
    for (int64_t i = 0; i < n; ++i) {
        volatile double prod = scores[i] * scale; /* volatile: forbid fused/extended eval */
        out[i] = (int64_t)prod;
    }
}

/* ---------------------------------------------------------------------------------
 * Term -> document inversion — src/deep_impact/inverted_index/create.py:31-51
 *
 * Input: a doc-major collection already mapped to term ids (create.py:19-29 gives
 * term_id = rank of the term string in sorted() order; that string sort is host work
 * done in Python by the caller). doc_offsets has n_docs+1 entries; postings of doc d are
 * term_ids/impacts[doc_offsets[d] .. doc_offsets[d+1]) in dict-insertion order.
 *
 * Output: term-major CSR. create.py:33-35 appends (doc_id, int(val)) to the term's list
 * in doc order; create.py:41 sorts each list with sorted(key=impact, reverse=True), which
 * is STABLE, so equal impacts keep ascending doc order.
 * ------------------------------------------------------------------------------- */
int dio_invert(const uint32_t *term_ids, const uint8_t *impacts, const uint64_t *doc_offsets,
               uint64_t n_docs, uint32_t n_terms,
               uint64_t *term_offsets /* n_terms+1 */, uint32_t *out_docids, uint8_t *out_impacts)
{
    uint64_t n_post = doc_offsets[n_docs];
    uint64_t *cursor = (uint64_t *)calloc((size_t)n_terms + 1, sizeof(uint64_t));
    uint32_t *tmp_doc = (uint32_t *)malloc((size_t)(n_post ? n_post : 1) * sizeof(uint32_t));
    uint8_t *tmp_imp = (uint8_t *)malloc((size_t)(n_post ? n_post : 1));
    if (!cursor || !tmp_doc || !tmp_imp) { free(cursor); free(tmp_doc); free(tmp_imp); return -1; }

    memset(term_offsets, 0, ((size_t)n_terms + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n_post; ++i) {
        if (term_ids[i] >= n_terms) { free(cursor); free(tmp_doc); free(tmp_imp); return -2; }
        term_offsets[term_ids[i] + 1]++;
    }
    for (uint32_t t = 0; t < n_terms; ++t) term_offsets[t + 1] += term_offsets[t];
    memcpy(cursor, term_offsets, ((size_t)n_terms + 1) * sizeof(uint64_t));

    /* create.py:33-35 — lists grow in document order */
    for (uint64_t d = 0; d < n_docs; ++d)
        for (uint64_t i = doc_offsets[d]; i < doc_offsets[d + 1]; ++i) {
            uint64_t p = cursor[term_ids[i]]++;
            tmp_doc[p] = (uint32_t)d;
            tmp_imp[p] = impacts[i];
        }

    /* create.py:41 — per-term stable sort by impact, descending. A stable counting sort
     * over the 256 possible impact values is the same permutation. */
    for (uint32_t t = 0; t < n_terms; ++t) {
        uint64_t lo = term_offsets[t], hi = term_offsets[t + 1];
        uint64_t cnt[256];
        memset(cnt, 0, sizeof cnt);
        for (uint64_t p = lo; p < hi; ++p) cnt[tmp_imp[p]]++;
        uint64_t pos[256], run = lo;
        for (int v = 255; v >= 0; --v) { pos[v] = run; run += cnt[v]; }
        for (uint64_t p = lo; p < hi; ++p) {
            uint64_t q = pos[tmp_imp[p]]++;
            out_docids[q] = tmp_doc[p];
            out_impacts[q] = tmp_imp[p];
        }
    }
    free(cursor); free(tmp_doc); free(tmp_imp);
    return 0;
}

/* create.py:44-51 + defaults.py:27-37 — serialise to the on-disk byte format:
 * .dat = per posting pack('I', doc) + pack('B', val) (5 bytes, little-endian, unpadded);
 * .idx = per term pack('Q', start_byte) + pack('Q', end_byte).
 * (A term with no postings cannot occur in the reference: every vocab term came from a doc.) */
void dio_serialize(const uint64_t *term_offsets, const uint32_t *docids, const uint8_t *impacts,
                   uint32_t n_terms, uint8_t *dat /* 5*P */, uint64_t *idx /* 2*n_terms */)
{
    uint64_t n_post = term_offsets[n_terms];
    for (uint64_t p = 0; p < n_post; ++p) {
        uint32_t d = docids[p];
        dat[5 * p + 0] = (uint8_t)(d);
        dat[5 * p + 1] = (uint8_t)(d >> 8);
        dat[5 * p + 2] = (uint8_t)(d >> 16);
        dat[5 * p + 3] = (uint8_t)(d >> 24);
        dat[5 * p + 4] = impacts[p];
    }
    for (uint32_t t = 0; t < n_terms; ++t) {
        idx[2 * t] = 5 * term_offsets[t];
        idx[2 * t + 1] = 5 * term_offsets[t + 1];
    }
}

/* ---------------------------------------------------------------------------------
 * Index reader — src/deep_impact/inverted_index/inverted_index.py:31-53
 * term_docs(): read 5-byte records from `start` while position < `end`, STOP at the
 * first record whose value == 0 (inverted_index.py:50-51). Returns the number of visible
 * postings; writes them to docs/vals when non-NULL.
 * ------------------------------------------------------------------------------- */
static inline uint32_t rd_u32le(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

int64_t dio_term_docs(const uint8_t *dat, uint64_t dat_bytes, uint64_t start, uint64_t end,
                      uint32_t *docs, uint8_t *vals)
{
    int64_t n = 0;
    for (uint64_t pos = start; pos < end; pos += 5) {
        if (pos + 5 > dat_bytes) return -1; /* struct.error in the reference: short read */
        uint8_t v = dat[pos + 4];
        if (v == 0) break;
        if (docs) docs[n] = rd_u32le(dat + pos);
        if (vals) vals[n] = v;
        ++n;
    }
    return n;
}

/* ---------------------------------------------------------------------------------
 * Scoring + top-k — src/deep_impact/inverted_index/inverted_index.py:55-62
 *
 * score(): for every query term IN THE ORDER GIVEN (duplicates count again), for every
 * visible posting: scores[doc] += impact (:58-60, a dict: docs are "touched" in first-seen
 * order). Then heapq.nlargest(top_k, items, key=score) (:62) == the first top_k items of a
 * STABLE sort by score descending, i.e. ties stay in first-touch order.
 *
 * tie_mode 0 ("raw")       : exactly that order (depends on query-term order).
 * tie_mode 1 ("canonical") : ties broken by ascending docid — the order-independent
 *                            definition the B200 path implements.
 *
 * The index is given as the raw file bytes (dat + idx pairs), so the reader semantics
 * above (break at value == 0) are part of what is checked.
 * Outputs: out_docs/out_scores hold up to top_k entries per query at stride top_k;
 * out_counts[q] = number of entries; out_postings[q] = postings traversed (for throughput).
 * ------------------------------------------------------------------------------- */
typedef struct { uint32_t doc; int32_t score; uint32_t touch; } dio_hit;

static int cmp_raw(const void *a, const void *b)
{
    const dio_hit *x = (const dio_hit *)a, *y = (const dio_hit *)b;
    if (x->score != y->score) return (x->score < y->score) ? 1 : -1;
    return (x->touch > y->touch) - (x->touch < y->touch);
}
static int cmp_canon(const void *a, const void *b)
{
    const dio_hit *x = (const dio_hit *)a, *y = (const dio_hit *)b;
    if (x->score != y->score) return (x->score < y->score) ? 1 : -1;
    return (x->doc > y->doc) - (x->doc < y->doc);
}

/* select the best k of hits[0..n) under cmp and sort them: nth-element style partition on
 * the (score, tiebreak) order, then qsort of the survivors. */
static void topk_sort(dio_hit *hits, int64_t n, int64_t k, int (*cmp)(const void *, const void *))
{
    if (k < n) {
        int64_t lo = 0, hi = n - 1;
        while (lo < hi) { /* quickselect so that hits[0..k) are the k best */
            dio_hit pivot = hits[lo + (hi - lo) / 2];
            int64_t i = lo, j = hi;
            while (i <= j) {
                while (cmp(&hits[i], &pivot) < 0) ++i;
                while (cmp(&hits[j], &pivot) > 0) --j;
                if (i <= j) { dio_hit t = hits[i]; hits[i] = hits[j]; hits[j] = t; ++i; --j; }
            }
            if (k - 1 <= j) hi = j; else if (k - 1 >= i) lo = i; else break;
        }
        n = k;
    }
    qsort(hits, (size_t)n, sizeof(dio_hit), cmp);
}

int dio_score_topk(const uint8_t *dat, uint64_t dat_bytes, const uint64_t *idx, uint32_t n_terms,
                   uint32_t n_docs,
                   const int64_t *q_term_ids /* -1 = term not in vocab */, const uint64_t *q_offsets,
                   uint32_t n_queries, uint32_t top_k, int tie_mode, int n_threads,
                   uint32_t *out_docs, int32_t *out_scores, uint32_t *out_counts,
                   uint64_t *out_postings)
{
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        int32_t *acc = (int32_t *)calloc((size_t)n_docs + 1, sizeof(int32_t));
        uint32_t *touched = (uint32_t *)malloc(((size_t)n_docs + 1) * sizeof(uint32_t));
        dio_hit *hits = (dio_hit *)malloc(((size_t)n_docs + 1) * sizeof(dio_hit));
        if (!acc || !touched || !hits) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t q = 0; q < (int64_t)n_queries; ++q) {
                uint64_t n_touched = 0, n_post = 0;
                int bad = 0;
                for (uint64_t j = q_offsets[q]; j < q_offsets[q + 1] && !bad; ++j) {
                    int64_t t = q_term_ids[j];
                    if (t < 0 || (uint64_t)t >= n_terms) continue; /* inverted_index.py:43-44 */
                    uint64_t start = idx[2 * t], end = idx[2 * t + 1];
                    for (uint64_t pos = start; pos < end; pos += 5) {
                        if (pos + 5 > dat_bytes) { bad = 1; break; }
                        uint8_t v = dat[pos + 4];
                        if (v == 0) break; /* inverted_index.py:50-51 */
                        uint32_t d = rd_u32le(dat + pos);
                        if (d >= n_docs) { bad = 1; break; }
                        if (acc[d] == 0) touched[n_touched++] = d; /* v >= 1, so 0 <=> untouched */
                        acc[d] += v;
                        ++n_post;
                    }
                }
                if (bad) {
#pragma omp atomic write
                    err = -2;
                }
                for (uint64_t i = 0; i < n_touched; ++i) {
                    hits[i].doc = touched[i];
                    hits[i].score = acc[touched[i]];
                    hits[i].touch = (uint32_t)i;
                    acc[touched[i]] = 0;
                }
                topk_sort(hits, (int64_t)n_touched, (int64_t)top_k, tie_mode ? cmp_canon : cmp_raw);
                uint64_t n_out = n_touched < top_k ? n_touched : top_k;
                for (uint64_t i = 0; i < n_out; ++i) {
                    out_docs[(uint64_t)q * top_k + i] = hits[i].doc;
                    out_scores[(uint64_t)q * top_k + i] = hits[i].score;
                }
                out_counts[q] = (uint32_t)n_out;
                if (out_postings) out_postings[q] = n_post;
            }
        }
        free(acc); free(touched); free(hits);
    }
    return err;
}

/* Same scorer over an in-memory CSR (term_offsets/docids/impacts, no zero impacts) — the
 * form the in-memory twin SparseSearch uses (evaluation/nano_beir_evaluator.py:103-137:
 * doc_scores[doc] += score per posting in list order, top-k by nlargest/sorted, stable).
 * Provided so large shards can be checked without materialising the 5-byte file image. */
int dio_score_topk_csr(const uint64_t *term_offsets, const uint32_t *docids, const uint8_t *impacts,
                       uint32_t n_terms, uint32_t n_docs,
                       const int64_t *q_term_ids, const uint64_t *q_offsets,
                       uint32_t n_queries, uint32_t top_k, int tie_mode, int n_threads,
                       uint32_t *out_docs, int32_t *out_scores, uint32_t *out_counts,
                       uint64_t *out_postings)
{
    int err = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        int32_t *acc = (int32_t *)calloc((size_t)n_docs + 1, sizeof(int32_t));
        uint32_t *touched = (uint32_t *)malloc(((size_t)n_docs + 1) * sizeof(uint32_t));
        dio_hit *hits = (dio_hit *)malloc(((size_t)n_docs + 1) * sizeof(dio_hit));
        if (!acc || !touched || !hits) {
#pragma omp atomic write
            err = -1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t q = 0; q < (int64_t)n_queries; ++q) {
                uint64_t n_touched = 0, n_post = 0;
                for (uint64_t j = q_offsets[q]; j < q_offsets[q + 1]; ++j) {
                    int64_t t = q_term_ids[j];
                    if (t < 0 || (uint64_t)t >= n_terms) continue;
                    for (uint64_t p = term_offsets[t]; p < term_offsets[t + 1]; ++p) {
                        uint8_t v = impacts[p];
                        if (v == 0) break;
                        uint32_t d = docids[p];
                        if (acc[d] == 0) touched[n_touched++] = d;
                        acc[d] += v;
                        ++n_post;
                    }
                }
                for (uint64_t i = 0; i < n_touched; ++i) {
                    hits[i].doc = touched[i];
                    hits[i].score = acc[touched[i]];
                    hits[i].touch = (uint32_t)i;
                    acc[touched[i]] = 0;
                }
                topk_sort(hits, (int64_t)n_touched, (int64_t)top_k, tie_mode ? cmp_canon : cmp_raw);
                uint64_t n_out = n_touched < top_k ? n_touched : top_k;
                for (uint64_t i = 0; i < n_out; ++i) {
                    out_docs[(uint64_t)q * top_k + i] = hits[i].doc;
                    out_scores[(uint64_t)q * top_k + i] = hits[i].score;
                }
                out_counts[q] = (uint32_t)n_out;
                if (out_postings) out_postings[q] = n_post;
            }
        }
        free(acc); free(touched); free(hits);
    }
    return err;
}

int dio_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
