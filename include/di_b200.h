/*
 * di_b200.h — C ABI of the B200-native DeeperImpact inverted-index path.
 *
 * The reference (Tommachilez/improving-learned-index) is pure Python and has no FFI; the
 * only callers of this ABI are the Python classes in improving-learned-index_b200/ that
 * mirror the reference's class surface. Each entry point names the reference code it
 * replaces (file:line relative to the reference root). INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types; every buffer is caller-owned; `*_dev` variants take DEVICE pointers and a
 *     cudaStream_t (passed as void*) and are asynchronous on that stream (the di_index_create_* constructors
 *     are the exception: they return a finished index); the others take HOST pointers and return when the
 *     result is in the output buffers.
 *   - return value: 0 = DI_OK, otherwise an error code; di_last_error() gives the message of
 *     the last failure on the calling thread.
 *   - no global mutable state besides the handles; a handle must not be used from two threads
 *     at once.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     DI_ERR_CUDA / DI_ERR_NODEVICE.
 */
#ifndef DI_B200_H
#define DI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DI_OK 0
#define DI_ERR_CUDA 1      /* a CUDA runtime call failed */
#define DI_ERR_ARG 2       /* invalid argument */
#define DI_ERR_RANGE 3     /* a value does not fit the index format (impact > 255, docid out of range, ...) */
#define DI_ERR_NOMEM 4     /* device or host allocation failed */
#define DI_ERR_NODEVICE 5  /* no CUDA device present */
#define DI_ERR_FORMAT 6    /* malformed index file image (short read) or collection line (a ValueError in the reference) */
#define DI_ERR_UNSUPPORTED 7 /* valid input outside the fast host parser's grammar: use the Python parser */

#define DI_OOV_TERM 0xFFFFFFFFu /* query term id meaning "not in vocabulary" (inverted_index.py:43-44) */

typedef struct di_index di_index_t; /* opaque, device-resident index shard */

const char *di_last_error(void);
int di_version(void);
int di_device_count(int *count);
int di_set_device(int device);

/* ------------------------------------------------------------------ K1: quantization
 * Replaces src/deep_impact/indexing/quantize.py:
 *   find_max_value :17-24  -> di_find_max_f64   (max seeded with 0.0)
 *   quantize       :13-14  -> di_quantize_f64   out[i] = (int)trunc(scores[i] * (255.0 / max_val)), float64
 * The caller applies quantize.py:45 (keep iff out[i] > 0). Results are saturated to the
 * int32 range, not to 255 (the reference does not clamp either).
 */
int di_find_max_f64(const double *scores, int64_t n, double *max_out);
int di_quantize_f64(const double *scores, int64_t n, double max_val, int32_t *out);
int di_find_max_f64_dev(const double *d_scores, int64_t n, double *d_max_out, void *stream);
int di_quantize_f64_dev(const double *d_scores, int64_t n, double max_val, int32_t *d_out, void *stream);

/* ------------------------------------------------------------------ collection text parser (host only, no GPU)
 * Replaces the Python string loops around K1/K2: DeepImpactCollection (deep_impact_collection.py:11-25),
 * the vocabulary of create.py:19-29 (term id = rank in sorted() order) and the parse/format halves of
 * quantize_file (quantize.py:17-24, 39-47). Text is the doc-major format "term: score, term: score",
 * one document per line. DI_PARSE_DICT = InvertedIndexCreator semantics (blank line = empty document, a
 * repeated term keeps its last score); DI_PARSE_SEQUENCE = quantize_file semantics (all pairs in order,
 * blank line = DI_ERR_FORMAT). The arrays returned by di_collection_arrays stay owned by the collection:
 * doc_offsets[n_docs+1], term_ids[n_postings] (ids = sorted rank), scores[n_postings] (float64), the
 * sorted vocabulary as one byte blob + vocab_offsets[n_terms+1].
 */
#define DI_PARSE_DICT 0
#define DI_PARSE_SEQUENCE 1
typedef struct di_collection di_collection_t;
int di_collection_parse(const char *text, uint64_t n_bytes, int mode, di_collection_t **out);
void di_collection_free(di_collection_t *collection);
int di_collection_info(const di_collection_t *collection, uint64_t *n_docs, uint64_t *n_postings, uint32_t *n_terms);
int di_collection_arrays(const di_collection_t *collection, const uint64_t **doc_offsets, const uint32_t **term_ids,
                         const double **scores, const char **vocab_blob, const uint64_t **vocab_offsets);
/* quantize.py:40-47: one line per document, "term: value" for every value > 0 joined by ", " */
int di_collection_write_quantized(const di_collection_t *collection, const int32_t *values, const char *path);

/* ------------------------------------------------------------------ K2: term -> document inversion
 * Replaces src/deep_impact/inverted_index/create.py:31-51 (InvertedIndexCreator._inverted_index).
 * Input is a doc-major collection already mapped to term ids (create.py:19-29: id = rank in
 * sorted() order; that string sort stays on the host): doc d owns postings
 * [doc_offsets[d], doc_offsets[d+1]).  Output is term-major CSR in the reference's order —
 * term ascending, impact descending, docid ascending (stable sort of create.py:41).
 *   term_offsets : n_terms + 1 entries;  out_docids / out_impacts : doc_offsets[n_docs] entries.
 * di_serialize produces the reference's on-disk images (create.py:44-51, utils/defaults.py:22-37):
 *   dat = 5-byte records pack('I', doc) + pack('B', impact);  idx = per term (start_byte, end_byte) as 2 x u64.
 */
int di_invert(const uint32_t *term_ids, const uint8_t *impacts, const uint64_t *doc_offsets,
              uint64_t n_docs, uint32_t n_terms,
              uint64_t *term_offsets, uint32_t *out_docids, uint8_t *out_impacts);
/* Device form: truly asynchronous on `stream` — scratch comes from the stream-ordered memory pool and nothing waits
 * for the GPU. *d_status (device u32, may be NULL) becomes 1 when a term id >= n_terms was met (di_invert returns
 * DI_ERR_RANGE for that); such postings are left out and d_term_offsets[n_terms] is the number kept. */
int di_invert_dev(const uint32_t *d_term_ids, const uint8_t *d_impacts, const uint64_t *d_doc_offsets,
                  uint64_t n_docs, uint32_t n_terms, uint64_t n_postings,
                  uint64_t *d_term_offsets, uint32_t *d_out_docids, uint8_t *d_out_impacts, uint32_t *d_status,
                  void *stream);
int di_serialize(const uint64_t *term_offsets, const uint32_t *docids, const uint8_t *impacts,
                 uint32_t n_terms, uint8_t *dat, uint64_t *idx);
int di_serialize_dev(const uint64_t *d_term_offsets, const uint32_t *d_docids, const uint8_t *d_impacts,
                     uint32_t n_terms, uint64_t n_postings, uint8_t *d_dat, uint64_t *d_idx, void *stream);

/* ------------------------------------------------------------------ index shard (device resident)
 * Replaces the reader half of src/deep_impact/inverted_index/inverted_index.py:24-53: instead of
 * re-opening .idx/.dat per term per query, the whole shard is re-laid out once in HBM as
 * document tiles. Reader semantics kept: within one term's list, postings at
 * or after the first impact == 0 are invisible (inverted_index.py:50-51).
 *
 * A shard holds the postings whose docid lies in [doc_lo, doc_hi); docids stay GLOBAL, so
 * results of different shards merge with the same deterministic order.
 */
typedef struct di_index_params {
    uint32_t tile_docs;      /* documents per tile: power of two in [256, 32768]; 0 = default (16384) */
    uint32_t dense_ratio;    /* a (term, tile) segment with n postings is stored as a dense array (one byte per
                                document of the tile) when n * dense_ratio >= tile_docs; 0 = default (8);
                                0xFFFFFFFF = never */
    uint32_t cand_slack;     /* per-query candidate slots kept between tiles; 0 = default (max(2k, 256)) */
    uint32_t flags;          /* DI_INDEX_* bits; 0 = default */
} di_index_params;
#define DI_INDEX_NO_SEEDS 1u /* build no threshold-seed tables (every query then starts without a bound) */
#define DI_INDEX_PER_TILE 2u /* diagnostic: one kernel launch per document tile instead of one persistent launch */
#define DI_INDEX_NO_BANK_SORT 8u /* A/B switch: keep sparse postings in docid order (default: runs of 128 postings are ordered by
                                    shared-memory bank, which makes the scorer's atomics conflict-free) */
#define DI_INDEX_TILE_BOUNDS 4u /* keep the largest impact of every (tile, term): a (query, tile) whose bounds add up to less
                                   than the query's running threshold is skipped without being scored — a proof (MaxScore
                                   style), results are unchanged. Pays when impacts are skewed across the document range
                                   (e.g. documents ordered by quality); costs one byte per (tile, term) and one byte load
                                   per query term and tile */

typedef struct di_index_info {
    uint64_t n_postings;      /* visible postings in the shard */
    uint64_t payload_bytes;   /* bytes of tiled posting payload in HBM */
    uint64_t table_bytes;     /* bytes of the (tile, term) segment table */
    uint64_t n_dense_segments;
    uint64_t n_sparse_segments;
    uint64_t n_dense_postings; /* postings held in dense segments */
    uint32_t n_terms;
    uint32_t doc_lo, doc_hi;
    uint32_t n_tiles, tile_docs;
    uint32_t max_docid_plus1; /* 1 + largest docid seen (0 if the shard is empty) */
} di_index_info;

/* from term-major CSR in host memory (any order inside a term's list is accepted) */
int di_index_create_csr(const uint64_t *term_offsets, const uint32_t *docids, const uint8_t *impacts,
                        uint32_t n_terms, uint32_t doc_lo, uint32_t doc_hi,
                        const di_index_params *params, di_index_t **out);
/* same, CSR already in device memory */
int di_index_create_csr_dev(const uint64_t *d_term_offsets, const uint32_t *d_docids, const uint8_t *d_impacts,
                            uint32_t n_terms, uint64_t n_postings, uint32_t doc_lo, uint32_t doc_hi,
                            const di_index_params *params, di_index_t **out);
/* straight from a DOC-MAJOR collection in device memory (the input of di_invert_dev: what create.py:31-35 iterates over),
 * for documents [doc_lo, doc_lo + n_docs): skips the term-major detour when no index files are wanted — one segmented
 * two-pass sort instead of inversion + re-tiling; the index is identical to the one built from the inverted CSR.
 * Postings with impact 0 are invisible (the reader never gets past them, inverted_index.py:50-51). */
int di_index_create_docmajor_dev(const uint32_t *d_term_ids, const uint8_t *d_impacts, const uint64_t *d_doc_offsets,
                                 uint64_t n_docs, uint32_t n_terms, uint64_t n_postings, uint32_t doc_lo,
                                 const di_index_params *params, di_index_t **out);
/* from the reference's file images: inverted_index.dat bytes + inverted_index.idx as (start,end) u64 pairs */
int di_index_create_files(const uint8_t *dat, uint64_t dat_bytes, const uint64_t *idx_pairs,
                          uint32_t n_terms, uint32_t doc_lo, uint32_t doc_hi,
                          const di_index_params *params, di_index_t **out);
void di_index_destroy(di_index_t *index);
int di_index_get_info(const di_index_t *index, di_index_info *info);
/* Seeds for a SHARD of a larger collection. A query's seed is a proven lower bound of its k-th best score over the documents
 * the seed tables count; with the tables of all shards added up, every shard starts from a bound of the GLOBAL k-th score and
 * does not emit what the cross-shard merge would discard. export: this shard's impact histogram of every term, dense
 * d_hist[n_terms][256] (u32, device). The caller adds the tables of all shards (e.g. one all-reduce) and imports the sum into
 * every shard. AFTER an import a search of this index returns only keys that can be in the top_k of the WHOLE collection: its
 * rows are meant for the merge, not a top_k of the shard on its own. (No-op for an index with duplicate postings.) */
int di_index_export_seed_hist_dev(const di_index_t *index, uint32_t *d_hist, void *stream);
int di_index_import_seed_hist_dev(di_index_t *index, const uint32_t *d_hist, void *stream);
/* Row order of di_search_dev's result (sticky; 0 = default: rows fully sorted). With sorted_prefix = p > 0 a row is
 * [its p best keys, sorted | the rest of its top_k keys in any order] — all a SHARD has to deliver to the cross-shard merge,
 * which reads the first k_in columns and, rarely, re-selects from the whole row; saves most of the per-row sort.
 * (Batches small enough to run in tile lanes are always fully sorted.) di_search (host rows) ignores it. */
int di_index_set_sorted_prefix(di_index_t *index, uint32_t sorted_prefix);
/* visible posting count (document frequency) of each given term id in this shard; OOV -> 0 */
int di_index_term_df(const di_index_t *index, const uint32_t *term_ids, uint64_t n, uint64_t *df_out);

/* ------------------------------------------------------------------ K3 + K4: scoring and top-k
 * Replaces InvertedIndex.score (inverted_index.py:55-62) for a BATCH of queries, and the
 * scoring loop of SparseSearch.search (evaluation/nano_beir_evaluator.py:113-133):
 *   score[q][doc] = sum over the query's term occurrences of the doc's impact (duplicates count
 *   again, inverted_index.py:58-60); result = the top_k docs with score > 0 ordered by score
 *   descending, TIES BY ASCENDING DOCID (the canonical order of SURVEY.md §8a; the reference's
 *   own tie order depends on PYTHONHASHSEED).
 * Query q owns term ids q_terms[q_offsets[q] .. q_offsets[q+1]); DI_OOV_TERM entries are skipped.
 * Outputs are row-major [n_queries][top_k]; rows are valid up to out_counts[q].
 * top_k <= 65536.
 * Scoring is exhaustive; a document is left out of a candidate list only when it is PROVEN to be outside the
 * top_k: below the running k-th best key, or below the query's seed bound (the k-th highest impact of one of
 * its frequent terms, from per-term impact tables built with the index; skipped for an index in which a posting
 * list names a document twice). The result is identical to scoring everything and sorting.
 */
int di_search(di_index_t *index, const uint32_t *q_terms, const uint64_t *q_offsets,
              uint32_t n_queries, uint32_t top_k,
              uint32_t *out_docids, int32_t *out_scores, uint32_t *out_counts);
/* device variant: packed keys (score << 32 | ~docid), sorted descending (or only up to di_index_set_sorted_prefix
 * columns), for the multi-GPU merge. Asynchronous on `stream`; the index's workspace serves one search at a time.
 * max_query_len = longest query in the batch (chooses 16- or 32-bit accumulators).
 * d_theta_init (optional, may be NULL): per query a key that the caller KNOWS to be a lower bound of the
 * query's final k-th best key over the whole collection (e.g. the k-th key of a previous, partial merge);
 * documents below it are not returned, which makes a second search round cheap. */
int di_search_dev(di_index_t *index, const uint32_t *d_q_terms, const uint64_t *d_q_offsets,
                  uint32_t n_queries, uint32_t max_query_len, uint32_t top_k, const uint64_t *d_theta_init,
                  uint64_t *d_out_keys, uint32_t *d_out_counts, void *stream);
int di_unpack_keys_dev(const uint64_t *d_keys, uint64_t n, uint32_t *d_docids, int32_t *d_scores, void *stream);

/* ------------------------------------------------------------------ K5: cross-shard top-k merge
 * New functionality (the reference is single-host): d_keys_in holds n_shards blocks of
 * [n_queries][k_in] sorted keys (as gathered over NCCL), d_counts_in n_shards x [n_queries].
 * Output: the global top_k per query in the same deterministic order, rows of top_k keys.
 * Shards may return fewer keys than top_k (k_in < top_k, less traffic and less per-shard selection
 * work). The merge is then exact unless a shard that filled its row could still hold a better key; if
 * d_incomplete is not NULL it receives 1 for exactly those queries (0 otherwise) and the caller re-runs
 * them with k_in = top_k (improving-learned-index_b200/sharded.py does). With k_in == top_k no query is
 * ever flagged.
 */
int di_merge_topk_dev(const uint64_t *d_keys_in, const uint32_t *d_counts_in,
                      uint32_t n_shards, uint32_t n_queries, uint32_t k_in, uint32_t top_k,
                      uint64_t *d_keys_out, uint32_t *d_counts_out, uint32_t *d_incomplete, void *stream);

/* K5 over peer memory: the exchange step as ONE fused kernel, no collective. d_rows / d_counts are DEVICE-resident tables
 * of n_shards pointers to every shard's sorted key rows ([.][row_stride]) and counts, where that shard's own
 * di_search_dev wrote them — peer GPU memory mapped into this process (di_shared_alloc / di_shared_open below); the
 * caller orders the shards' searches before this call with di_peer_barrier_dev. The queries are partitioned over the
 * ranks: this call merges queries [q_first, q_first + n_queries) only; output rows / counts are indexed from 0.
 * Per query, one CTA pulls the first min(count, k_in) keys of each shard's row over NVLink into shared memory, selects
 * and sorts the top_k, proves the result complete (a shard holding more than k_in keys hides only keys below its
 * k_in-th) and, for a query that fails the proof, pulls the full rows and selects again (a second launch of the same
 * shape in which the CTAs of proven queries leave at once) — no host round trip, no compaction of query ids.
 * Rows need to be sorted only in their first k_in columns. *d_n_second_pass (optional counter, caller-zeroed) counts the
 * queries of the second pass. row_stride <= top_k. */
int di_merge_pull_dev(const uint64_t *const *d_rows, const uint32_t *const *d_counts, uint32_t n_shards,
                      uint32_t q_first, uint32_t n_queries, uint32_t row_stride, uint32_t k_in, uint32_t top_k,
                      uint64_t *d_keys_out, uint32_t *d_counts_out, uint32_t *d_n_second_pass, void *stream);

/* Peer-visible device memory for the rows above (one process per GPU): di_shared_alloc = cudaMalloc (zero-filled) +
 * a 64-byte CUDA IPC handle that the caller sends to the other ranks (any transport); di_shared_open maps a peer's
 * allocation into this process (peer access over NVLink is enabled on first use). */
int di_shared_alloc(uint64_t bytes, void **d_ptr, uint8_t handle_out[64]);
int di_shared_open(const uint8_t handle[64], void **d_ptr);
int di_shared_close(void *d_ptr);
int di_shared_free(void *d_ptr);
/* Cross-GPU barrier ON THE STREAM (a one-CTA kernel, no host involvement): d_flags is a device table of n_ranks
 * pointers, entry r = rank r's flag array of n_ranks u32 (di_shared_alloc'ed, zero at start). Every rank calls it
 * with the same, strictly increasing epoch (1, 2, ...). Work enqueued before it on every rank's stream is complete
 * and visible to all ranks' work enqueued after it. */
int di_peer_barrier_dev(uint32_t *const *d_flags, uint32_t n_ranks, uint32_t my_rank, uint32_t epoch, void *stream);

/* page-locked host memory for query / result buffers of di_search (optional: any host memory works, pinned memory makes
 * the copies run at PCIe rate) */
int di_host_alloc(uint64_t bytes, void **ptr);
int di_host_free(void *ptr);

/* ------------------------------------------------------------------ after the search: run file and metrics
 * di_write_run_file replaces RunFile.writelines (src/utils/datasets.py:312-317) for a whole batch: appends the rows
 * "qid<TAB>pid<TAB>rank<TAB>score\n" (rank from 1) of n_queries result lists to `path`, formatted and written by all
 * host threads straight from the result arrays (host pointers: rows of row_stride entries, counts[i] valid). Query i's
 * id is the bytes qid_blob[qid_offsets[i] .. qid_offsets[i+1]). The bytes are those the reference writer appends.
 *
 * di_eval_ranks_dev replaces the run-file pass of Metrics.evaluate (src/deep_impact/evaluation/metrics.py:31-43) while
 * the result keys are still on the device: query i's relevant documents are the SORTED docids
 * d_qrel_docs[d_qrel_offsets[i] .. d_qrel_offsets[i+1]); d_best_rank[i] = rank (from 1) of its first relevant hit (0 =
 * none retrieved), d_hits[i * n_depths + j] = relevant hits with rank <= d_depths[j] (n_depths <= 8). The float
 * arithmetic of metrics.py:36-43 stays with the caller, so the reported numbers are bit-identical. */
/* The same as a stream: submit() formats a batch on the calling thread (all host threads) and returns once the PREVIOUS
 * batch has reached the page cache — the copy of a batch into the file overlaps the formatting (and the search) of the next
 * one, and the caller's arrays are free again when submit() returns. Batches reach the file in submission order. */
typedef struct di_run_writer di_run_writer_t;
int di_run_writer_open(const char *path, di_run_writer_t **out);
int di_run_writer_submit(di_run_writer_t *writer, const char *qid_blob, const uint64_t *qid_offsets, const uint32_t *docids,
                         const int32_t *scores, const uint32_t *counts, uint32_t n_queries, uint32_t row_stride);
int di_run_writer_close(di_run_writer_t *writer);
int di_write_run_file(const char *path, const char *qid_blob, const uint64_t *qid_offsets, const uint32_t *docids,
                      const int32_t *scores, const uint32_t *counts, uint32_t n_queries, uint32_t row_stride);
int di_eval_ranks_dev(const uint64_t *d_keys, const uint32_t *d_counts, uint32_t n_queries, uint32_t row_stride,
                      const uint64_t *d_qrel_offsets, const uint32_t *d_qrel_docs, const uint32_t *d_depths,
                      uint32_t n_depths, uint32_t *d_best_rank, uint32_t *d_hits, void *stream);

/* ------------------------------------------------------------------ measurement hooks
 * Device time (CUDA events on the launching stream) of the di_search / di_search_dev calls made since the
 * previous di_get_timings() (reading clears the record). */
typedef struct di_timings {
    float score_ms;      /* score kernel launches */
    float finalize_ms;   /* final select + sort */
    float total_ms;      /* first kernel to last kernel */
    uint32_t score_launches;
    uint32_t other_launches;
    uint32_t lanes;      /* tile lanes per query of the last search (1 = one chain of tiles per query) */
    uint32_t acc32;      /* 1 when the last search ran with 32-bit accumulators */
    uint64_t tiles_skipped; /* (query, tile) pairs proven empty of results by DI_INDEX_TILE_BOUNDS and not scored */
} di_timings;
int di_get_timings(di_index_t *index, di_timings *out);

#ifdef __cplusplus
}
#endif
#endif /* DI_B200_H */
