// di_b200.cu — the C ABI declared in include/di_b200.h. Host-side orchestration only; the
// kernels live in build.cuh / search.cuh / scan_sort.cuh. Compiled for sm_100a.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <thread>
#include <vector>

#include "build.cuh"
#include "common.cuh"
#include "scan_sort.cuh"
#include "search.cuh"

using namespace di;

// ============================================================================ handle
struct di_index {
    int device = 0;
    uint32_t n_terms = 0, doc_lo = 0, doc_hi = 0;
    uint32_t n_tiles = 0, tile_docs = 0, tile_shift = 0, max_docid_plus1 = 0;
    uint32_t dense_ratio = 8, cand_slack = 0, flags = 0;
    bool has_dup_postings = false;  // some posting list names a document twice: no seeds, 32-bit accumulators
    uint32_t last_lanes = 0, last_acc32 = 0;
    uint32_t sorted_prefix = 0;   // di_index_set_sorted_prefix: 0 = result rows fully sorted
    bool global_seeds = false;    // the seed tables count the postings of a larger collection (di_index_import_seed_hist_dev)
    uint64_t n_postings = 0, payload_bytes = 0, table_bytes = 0;
    uint64_t n_dense_segments = 0, n_sparse_segments = 0, n_dense_postings = 0;
    SegDesc *d_desc = nullptr;
    uint8_t *d_payload = nullptr;
    unsigned long long *d_df = nullptr;
    // threshold seeds (build.cuh): per-term tables cum[v] = #postings with impact >= v for the frequent terms
    uint32_t *d_seed_slot = nullptr, *d_seed_cum = nullptr;
    uint8_t *d_seg_max = nullptr;   // [n_tiles][n_terms] largest impact per segment (DI_INDEX_TILE_BOUNDS), else nullptr
    DevBuf ws_skipped;
    uint64_t last_skipped = 0;

    // search workspace (grown on demand)
    cudaStream_t stream = nullptr;
    DevBuf ws_cand, ws_cnt, ws_theta, ws_order, ws_done, ws_lane_keys, ws_lane_counts;
    DevBuf st_qterms, st_qoffs, st_keys, st_counts, st_docids, st_scores;
    int smem_opt_in = 0;
    uint32_t attr_set = 0;   // score kernel variants whose function attributes are set on this device

    // timing of the last search
    static constexpr int kMaxBatches = 64;
    cudaEvent_t ev[kMaxBatches][3] = {};
    int n_batches = 0;
    uint32_t score_launches = 0, other_launches = 0;

    ~di_index()
    {
        if (d_desc) cudaFree(d_desc);
        if (d_payload) cudaFree(d_payload);
        if (d_df) cudaFree(d_df);
        if (d_seed_slot) cudaFree(d_seed_slot);
        if (d_seed_cum) cudaFree(d_seed_cum);
        if (d_seg_max) cudaFree(d_seg_max);
        for (auto &b : ev)
            for (auto &e : b)
                if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_error(DI_ERR_NODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    return DI_OK;
}

extern "C" const char *di_last_error(void) { return err_buf(); }
extern "C" int di_version(void) { return 100; }

extern "C" int di_device_count(int *count)
{
    if (!count) return set_error(DI_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return DI_OK;
}

extern "C" int di_set_device(int device)
{
    DI_TRY(ensure_device());
    DI_CUDA(cudaSetDevice(device));
    return DI_OK;
}

// ============================================================================ K1
extern "C" int di_find_max_f64_dev(const double *d_scores, int64_t n, double *d_max_out, void *stream)
{
    if (n < 0 || !d_max_out) return set_error(DI_ERR_ARG, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    DI_CUDA(cudaMemsetAsync(d_max_out, 0, sizeof(double), st));  // bits of +0.0
    if (n) {
        max_f64_kernel<<<grid_for((uint64_t)n, 256, 148 * 8), 256, 0, st>>>(d_scores, n, (unsigned long long *)d_max_out);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

extern "C" int di_quantize_f64_dev(const double *d_scores, int64_t n, double max_val, int32_t *d_out, void *stream)
{
    if (n < 0) return set_error(DI_ERR_ARG, "n < 0");
    if (n == 0) return DI_OK;
    const double scale = 255.0 / max_val;  // quantize.py:37, float64 division on the host
    quantize_f64_kernel<<<grid_for((uint64_t)n, 256), 256, 0, (cudaStream_t)stream>>>(d_scores, n, scale, d_out);
    DI_KERNEL_CHECK();
    return DI_OK;
}

extern "C" int di_find_max_f64(const double *scores, int64_t n, double *max_out)
{
    if (n < 0 || !max_out) return set_error(DI_ERR_ARG, "bad arguments");
    DI_TRY(ensure_device());
    DevBuf d_x, d_m;
    DI_TRY(d_x.alloc((size_t)n * sizeof(double)));
    DI_TRY(d_m.alloc(sizeof(double)));
    if (n) DI_CUDA(cudaMemcpy(d_x.p, scores, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    DI_TRY(di_find_max_f64_dev(d_x.as<double>(), n, d_m.as<double>(), nullptr));
    DI_CUDA(cudaMemcpy(max_out, d_m.p, sizeof(double), cudaMemcpyDeviceToHost));
    return DI_OK;
}

extern "C" int di_quantize_f64(const double *scores, int64_t n, double max_val, int32_t *out)
{
    if (n < 0) return set_error(DI_ERR_ARG, "n < 0");
    DI_TRY(ensure_device());
    if (n == 0) return DI_OK;
    DevBuf d_x, d_o;
    DI_TRY(d_x.alloc((size_t)n * sizeof(double)));
    DI_TRY(d_o.alloc((size_t)n * sizeof(int32_t)));
    DI_CUDA(cudaMemcpy(d_x.p, scores, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    DI_TRY(di_quantize_f64_dev(d_x.as<double>(), n, max_val, d_o.as<int32_t>(), nullptr));
    DI_CUDA(cudaMemcpy(out, d_o.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return DI_OK;
}

// ============================================================================ K2
extern "C" int di_invert_dev(const uint32_t *d_term_ids, const uint8_t *d_impacts, const uint64_t *d_doc_offsets,
                             uint64_t n_docs, uint32_t n_terms, uint64_t n_postings, uint64_t *d_term_offsets,
                             uint32_t *d_out_docids, uint8_t *d_out_impacts, uint32_t *d_status, void *stream)
{
    return invert_dev(d_term_ids, d_impacts, d_doc_offsets, n_docs, n_terms, n_postings, d_term_offsets, d_out_docids,
                      d_out_impacts, d_status, (cudaStream_t)stream);
}

extern "C" int di_invert(const uint32_t *term_ids, const uint8_t *impacts, const uint64_t *doc_offsets, uint64_t n_docs,
                         uint32_t n_terms, uint64_t *term_offsets, uint32_t *out_docids, uint8_t *out_impacts)
{
    if (!doc_offsets || !term_offsets) return set_error(DI_ERR_ARG, "NULL argument");
    DI_TRY(ensure_device());
    const uint64_t P = doc_offsets[n_docs];
    DevBuf d_t, d_v, d_o, d_to, d_od, d_ov, d_status;
    DI_TRY(d_status.alloc(sizeof(uint32_t)));
    DI_TRY(d_t.alloc(P * sizeof(uint32_t)));
    DI_TRY(d_v.alloc(P));
    DI_TRY(d_o.alloc((n_docs + 1) * sizeof(uint64_t)));
    DI_TRY(d_to.alloc(((size_t)n_terms + 1) * sizeof(uint64_t)));
    DI_TRY(d_od.alloc(P * sizeof(uint32_t)));
    DI_TRY(d_ov.alloc(P));
    if (P) {
        DI_CUDA(cudaMemcpy(d_t.p, term_ids, P * sizeof(uint32_t), cudaMemcpyHostToDevice));
        DI_CUDA(cudaMemcpy(d_v.p, impacts, P, cudaMemcpyHostToDevice));
    }
    DI_CUDA(cudaMemcpy(d_o.p, doc_offsets, (n_docs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    DI_TRY(invert_dev(d_t.as<uint32_t>(), d_v.as<uint8_t>(), d_o.as<uint64_t>(), n_docs, n_terms, P, d_to.as<uint64_t>(),
                      d_od.as<uint32_t>(), d_ov.as<uint8_t>(), d_status.as<uint32_t>(), nullptr));
    uint32_t status = 0;
    DI_CUDA(cudaMemcpy(&status, d_status.p, sizeof status, cudaMemcpyDeviceToHost));
    if (status) return set_error(DI_ERR_RANGE, "term id >= n_terms (%u) in the collection", n_terms);
    DI_CUDA(cudaMemcpy(term_offsets, d_to.p, ((size_t)n_terms + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (P) {
        DI_CUDA(cudaMemcpy(out_docids, d_od.p, P * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        DI_CUDA(cudaMemcpy(out_impacts, d_ov.p, P, cudaMemcpyDeviceToHost));
    }
    return DI_OK;
}

extern "C" int di_serialize_dev(const uint64_t *d_term_offsets, const uint32_t *d_docids, const uint8_t *d_impacts,
                                uint32_t n_terms, uint64_t n_postings, uint8_t *d_dat, uint64_t *d_idx, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n_postings) {
        serialize_dat_kernel<<<grid_for(n_postings, 256), 256, 0, st>>>(d_docids, d_impacts, n_postings, d_dat);
        DI_KERNEL_CHECK();
    }
    if (n_terms) {
        serialize_idx_kernel<<<grid_for(n_terms, 256), 256, 0, st>>>(d_term_offsets, n_terms, d_idx);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

extern "C" int di_serialize(const uint64_t *term_offsets, const uint32_t *docids, const uint8_t *impacts, uint32_t n_terms,
                            uint8_t *dat, uint64_t *idx)
{
    if (!term_offsets) return set_error(DI_ERR_ARG, "NULL argument");
    DI_TRY(ensure_device());
    const uint64_t P = term_offsets[n_terms];
    DevBuf d_to, d_d, d_v, d_dat, d_idx;
    DI_TRY(d_to.alloc(((size_t)n_terms + 1) * sizeof(uint64_t)));
    DI_TRY(d_d.alloc(P * sizeof(uint32_t)));
    DI_TRY(d_v.alloc(P));
    DI_TRY(d_dat.alloc(P * 5));
    DI_TRY(d_idx.alloc((size_t)n_terms * 16));
    DI_CUDA(cudaMemcpy(d_to.p, term_offsets, ((size_t)n_terms + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    if (P) {
        DI_CUDA(cudaMemcpy(d_d.p, docids, P * sizeof(uint32_t), cudaMemcpyHostToDevice));
        DI_CUDA(cudaMemcpy(d_v.p, impacts, P, cudaMemcpyHostToDevice));
    }
    DI_TRY(di_serialize_dev(d_to.as<uint64_t>(), d_d.as<uint32_t>(), d_v.as<uint8_t>(), n_terms, P, d_dat.as<uint8_t>(),
                            d_idx.as<uint64_t>(), nullptr));
    if (P) DI_CUDA(cudaMemcpy(dat, d_dat.p, P * 5, cudaMemcpyDeviceToHost));
    if (n_terms) DI_CUDA(cudaMemcpy(idx, d_idx.p, (size_t)n_terms * 16, cudaMemcpyDeviceToHost));
    return DI_OK;
}

// ============================================================================ index create
// Second half of both tiled builds: keys sorted by (tile, term, parity, local) inside rs_result(ka, kb, cur)[0 .. n_keys)
// (hidden keys anywhere, recognised by their term field) -> segment table + payload. count_seeds: also count the
// impact histograms of the frequent terms from the keys (the term-major build has them already).
static int finish_tiled(di_index *ix, const uint64_t *ka, const uint64_t *kb, const uint32_t *cur, uint64_t n_keys,
                        TileStats *d_stats, bool count_seeds, cudaStream_t st)
{
    const uint32_t V = ix->n_terms;
    const uint64_t n_segs = (uint64_t)ix->n_tiles * V;
    if (n_segs * sizeof(SegDesc) > (48ull << 30))
        return set_error(DI_ERR_NOMEM, "segment table of %llu entries is too large; use larger tiles",
                         (unsigned long long)n_segs);
    TileStats stats{};
    // scratch from the stream-ordered pool: a rebuild (or a build after an inversion) reuses the blocks already mapped
    StreamBuf d_begin(st), d_end(st), d_odd(st), d_size(st), d_nflag(st), d_scan(st), d_cnt(st);
    DI_TRY(d_begin.alloc(n_segs * 4));
    DI_TRY(d_end.alloc(n_segs * 4));
    DI_TRY(d_odd.alloc(n_segs * 4));
    DI_CUDA(cudaMemsetAsync(d_odd.p, 0xFF, n_segs * 4, st));
    DI_TRY(d_size.alloc((n_segs + 1) * 4));
    DI_TRY(d_nflag.alloc(n_segs * 4));
    DI_TRY(d_scan.alloc(scan_scratch_words(n_segs + 1) * 4));
    DI_CUDA(cudaMalloc(&ix->d_df, (size_t)V * sizeof(unsigned long long)));
    DI_CUDA(cudaMemsetAsync(ix->d_df, 0, (size_t)V * sizeof(unsigned long long), st));
    DI_CUDA(cudaMemsetAsync(d_begin.p, 0, n_segs * 4, st));
    DI_CUDA(cudaMemsetAsync(d_end.p, 0, n_segs * 4, st));
    DI_CUDA(cudaMemsetAsync(d_size.p, 0, (n_segs + 1) * 4, st));
    seg_bounds_kernel<<<grid_for(n_keys, 256), 256, 0, st>>>(ka, kb, cur, n_keys, V, d_begin.as<uint32_t>(),
                                                           d_end.as<uint32_t>(), d_odd.as<uint32_t>(), d_size.as<uint32_t>());
    DI_KERNEL_CHECK();
    seg_size_kernel<<<grid_for(n_segs, 256), 256, 0, st>>>(d_begin.as<uint32_t>(), d_end.as<uint32_t>(),
                                                          d_odd.as<uint32_t>(), n_segs, V,
                                                          ix->tile_docs, ix->dense_ratio, d_size.as<uint32_t>(),
                                                          d_nflag.as<uint32_t>(), ix->d_df, d_stats);
    DI_KERNEL_CHECK();
    // exclusive scan over n_segs + 1 entries: the last output is the total payload size
    DI_TRY(exclusive_scan_u32(d_size.as<uint32_t>(), d_size.as<uint32_t>(), n_segs + 1, d_scan.as<uint32_t>(), st));
    uint32_t total16 = 0;
    DI_CUDA(cudaMemcpyAsync(&total16, d_size.as<uint32_t>() + n_segs, 4, cudaMemcpyDeviceToHost, st));
    DI_CUDA(cudaMemcpyAsync(&stats, d_stats, sizeof stats, cudaMemcpyDeviceToHost, st));
    const uint64_t max_slots = std::min<uint64_t>(V, n_keys / kSeedMinDf);
    const bool seeds = count_seeds && max_slots && !(ix->flags & DI_INDEX_NO_SEEDS);
    if (seeds) {  // which terms are frequent is known now (df): slots, then the histograms ride along with the payload fill
        DI_TRY(d_cnt.alloc(4));
        DI_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4, st));
        DI_CUDA(cudaMalloc(&ix->d_seed_slot, (size_t)V * 4));
        DI_CUDA(cudaMalloc(&ix->d_seed_cum, max_slots * 256 * 4));
        DI_CUDA(cudaMemsetAsync(ix->d_seed_cum, 0, max_slots * 256 * 4, st));
        seed_slots_from_df_kernel<<<grid_for(V, 256), 256, 0, st>>>(ix->d_df, V, ix->d_seed_slot, d_cnt.as<uint32_t>());
        DI_KERNEL_CHECK();
    }
    DI_CUDA(cudaStreamSynchronize(st));
    ix->has_dup_postings = stats.n_dup_segments != 0;
    ix->n_dense_segments = stats.n_dense_segments;
    ix->n_sparse_segments = stats.n_sparse_segments;
    ix->n_dense_postings = stats.n_dense_postings;
    ix->payload_bytes = (uint64_t)total16 * 16;
    ix->table_bytes = n_segs * sizeof(SegDesc);
    DI_CUDA(cudaMalloc(&ix->d_desc, ix->table_bytes));
    DI_CUDA(cudaMalloc(&ix->d_payload, ix->payload_bytes ? ix->payload_bytes : 16));
    DI_CUDA(cudaMemsetAsync(ix->d_payload, 0, ix->payload_bytes ? ix->payload_bytes : 16, st));
    seg_desc_kernel<<<grid_for(n_segs, 256), 256, 0, st>>>(d_size.as<uint32_t>(), d_nflag.as<uint32_t>(), n_segs, ix->d_desc);
    DI_KERNEL_CHECK();
    const bool bounds = (ix->flags & DI_INDEX_TILE_BOUNDS) && !ix->has_dup_postings;
    if (bounds) {  // d_end is free now: it collects the per-segment maxima as u32
        DI_CUDA(cudaMemsetAsync(d_end.p, 0, n_segs * 4, st));
        DI_CUDA(cudaMalloc(&ix->d_seg_max, n_segs));
    }
    fill_payload_kernel<<<grid_for(n_keys, 256), 256, 0, st>>>(ka, kb, cur, n_keys, V, ix->d_desc, d_begin.as<uint32_t>(),
                                                               d_odd.as<uint32_t>(), ix->tile_docs, ix->d_payload,
                                                               seeds ? ix->d_seed_slot : nullptr, seeds ? ix->d_seed_cum : nullptr,
                                                               bounds ? d_end.as<uint32_t>() : nullptr);
    DI_KERNEL_CHECK();
    if (bounds) {
        narrow_u8_kernel<<<grid_for(n_segs, 256), 256, 0, st>>>(d_end.as<uint32_t>(), n_segs, ix->d_seg_max);
        DI_KERNEL_CHECK();
    }
    if (!(ix->flags & DI_INDEX_NO_BANK_SORT)) {
        sparse_bank_sort_kernel<<<grid_for(n_segs, kBankSortWarps, 148 * 64), kBankSortWarps * 32, 0, st>>>(ix->d_desc, n_segs,
                                                                                                         ix->d_payload);
        DI_KERNEL_CHECK();
    }
    if (seeds) {
        seed_cum_kernel<<<(unsigned)((max_slots * 32 + 255) / 256), 256, 0, st>>>(ix->d_seed_cum, (uint32_t)max_slots);
        DI_KERNEL_CHECK();
    }
    DI_CUDA(cudaStreamSynchronize(st));
    if (ix->has_dup_postings && ix->d_seed_cum) {  // a posting list that names a document twice: k postings != k documents
        cudaFree(ix->d_seed_cum);
        cudaFree(ix->d_seed_slot);
        ix->d_seed_cum = ix->d_seed_slot = nullptr;
    }
    return DI_OK;
}

// From term-major CSR (the reference's index format): postings are re-keyed and sorted by (tile, term, parity, local).
static int build_tiled(di_index *ix, const uint64_t *d_term_offsets, const uint32_t *d_docids, const uint8_t *d_impacts,
                       uint64_t n_post, cudaStream_t st)
{
    const uint32_t V = ix->n_terms;
    if (V >= kMaxTerms) return set_error(DI_ERR_ARG, "n_terms %u exceeds 2^24 - 1", V);
    if (n_post >= (1ull << 32) - 1) return set_error(DI_ERR_ARG, "more than 2^32-2 postings in one shard");

    DevBuf d_stats, d_fz;
    StreamBuf ka(st), kb(st);
    RadixSortScratch ws(st);
    DI_TRY(d_stats.alloc(sizeof(TileStats)));
    DI_CUDA(cudaMemsetAsync(d_stats.p, 0, sizeof(TileStats), st));
    TileStats stats{};
    if (n_post) {
        DI_TRY(d_fz.alloc((size_t)V * sizeof(unsigned long long)));
        DI_TRY(ka.alloc(n_post * sizeof(uint64_t)));
        DI_TRY(kb.alloc(n_post * sizeof(uint64_t)));
        init_first_zero_kernel<<<grid_for(V, 256), 256, 0, st>>>(d_term_offsets, V, d_fz.as<unsigned long long>());
        DI_KERNEL_CHECK();
        first_zero_kernel<<<grid_for(n_post, 256), 256, 0, st>>>(d_term_offsets, V, d_impacts, n_post,
                                                                 d_fz.as<unsigned long long>());
        DI_KERNEL_CHECK();
        tile_keys_kernel<<<grid_for(n_post, 256), 256, 0, st>>>(d_term_offsets, V, d_docids, d_impacts, n_post,
                                                                d_fz.as<unsigned long long>(), ix->doc_lo, ix->doc_hi,
                                                                (int)ix->tile_shift, ka.as<uint64_t>(),
                                                                d_stats.as<TileStats>());
        DI_KERNEL_CHECK();
        DI_CUDA(cudaMemcpyAsync(&stats, d_stats.p, sizeof stats, cudaMemcpyDeviceToHost, st));
        DI_CUDA(cudaStreamSynchronize(st));
        if (stats.bad_docid)
            return set_error(DI_ERR_RANGE, "shard spans more than 65535 tiles of %u docs; raise tile_docs or shard further",
                             ix->tile_docs);
        // threshold seeds: impact histograms of the frequent terms over the same visible postings
        const uint64_t max_slots = std::min<uint64_t>(V, n_post / kSeedMinDf);
        if (max_slots && !(ix->flags & DI_INDEX_NO_SEEDS)) {
            DevBuf d_cnt;
            DI_TRY(d_cnt.alloc(4));
            DI_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4, st));
            DI_CUDA(cudaMalloc(&ix->d_seed_slot, (size_t)V * 4));
            DI_CUDA(cudaMalloc(&ix->d_seed_cum, max_slots * 256 * 4));
            DI_CUDA(cudaMemsetAsync(ix->d_seed_cum, 0, max_slots * 256 * 4, st));
            seed_slots_kernel<<<grid_for(V, 256), 256, 0, st>>>(d_term_offsets, V, ix->d_seed_slot, d_cnt.as<uint32_t>());
            DI_KERNEL_CHECK();
            impact_hist_kernel<<<(unsigned)((n_post + kSeedChunk - 1) / kSeedChunk), 256, 0, st>>>(
                d_term_offsets, V, d_docids, d_impacts, n_post, d_fz.as<unsigned long long>(), ix->doc_lo, ix->doc_hi,
                ix->d_seed_slot, ix->d_seed_cum);
            DI_KERNEL_CHECK();
            seed_cum_kernel<<<(unsigned)((max_slots * 32 + 255) / 256), 256, 0, st>>>(ix->d_seed_cum, (uint32_t)max_slots);
            DI_KERNEL_CHECK();
            DI_CUDA(cudaStreamSynchronize(st));  // d_cnt goes out of scope
        }
        d_fz.release();
        // hidden postings carry ~0 and sort to the end; the (tile, term, local) order is total
        DI_TRY(radix_sort_u64(ka.as<uint64_t>(), kb.as<uint64_t>(), n_post, kTkLocalShift, 64, nullptr, 1, ws, st));
    }
    const uint64_t n_vis = stats.n_visible;
    ix->n_postings = n_vis;
    ix->max_docid_plus1 = stats.max_docid_plus1;
    ix->n_tiles = n_vis ? ((stats.max_docid_plus1 - ix->doc_lo) + ix->tile_docs - 1) / ix->tile_docs : 0;
    if (ix->n_tiles == 0) return DI_OK;
    return finish_tiled(ix, ka.as<uint64_t>(), kb.as<uint64_t>(), ws.cur(), n_vis, d_stats.as<TileStats>(), false, st);
}

// From a doc-major collection (what the indexer produces, create.py:31-35): the input is already grouped by tile and
// ordered by document inside a tile, so ONE segmented stable sort of every tile on (term, parity) — two 8-bit
// passes for a BERT-sized vocabulary, working set of a tile ~ L2 — replaces the inversion (3 passes) plus the
// re-tiling sort (6 passes) of the term-major route. The resulting index is identical.
static int build_tiled_docmajor(di_index *ix, const uint32_t *d_term_ids, const uint8_t *d_impacts,
                                const uint64_t *d_doc_offsets, uint64_t n_docs, uint64_t n_post, cudaStream_t st)
{
    const uint32_t V = ix->n_terms;
    if (V >= kMaxTerms) return set_error(DI_ERR_ARG, "n_terms %u exceeds 2^24 - 1", V);
    if (n_post >= (1ull << 32) - 1) return set_error(DI_ERR_ARG, "more than 2^32-2 postings in one shard");
    const uint64_t n_tiles = (n_docs + ix->tile_docs - 1) / ix->tile_docs;
    if (n_tiles >= 0xFFFFu)
        return set_error(DI_ERR_RANGE, "shard spans more than 65535 tiles of %u docs; raise tile_docs or shard further", ix->tile_docs);
    DevBuf d_stats;
    StreamBuf ka(st), kb(st), d_first(st);
    RadixSortScratch ws(st);
    DI_TRY(d_stats.alloc(sizeof(TileStats)));
    DI_CUDA(cudaMemsetAsync(d_stats.p, 0, sizeof(TileStats), st));
    ix->n_tiles = 0;
    if (n_post == 0 || n_docs == 0) return DI_OK;
    DI_TRY(ka.alloc(n_post * sizeof(uint64_t)));
    DI_TRY(kb.alloc(n_post * sizeof(uint64_t)));
    DI_TRY(d_first.alloc((n_tiles + 1) * sizeof(uint64_t)));
    tile_keys_docmajor_kernel<<<grid_for(n_post, 256), 256, 0, st>>>(d_term_ids, d_impacts, d_doc_offsets, n_docs, V,
                                                                     (int)ix->tile_shift, ka.as<uint64_t>(),
                                                                     d_stats.as<TileStats>());
    DI_KERNEL_CHECK();
    tile_first_keys_kernel<<<grid_for(n_tiles + 1, 256), 256, 0, st>>>(d_doc_offsets, n_docs, (int)ix->tile_shift,
                                                                       (uint32_t)n_tiles, d_first.as<uint64_t>());
    DI_KERNEL_CHECK();
    int term_bits = 1;
    while ((1ull << term_bits) < (uint64_t)V + 1) ++term_bits;  // the value V marks hidden postings
    // bit 23 = parity of the document, bits 24.. = term: inside a tile the input is in document order already
    DI_TRY(radix_sort_u64(ka.as<uint64_t>(), kb.as<uint64_t>(), n_post, kTkTermShift - 1, kTkTermShift + term_bits,
                          d_first.as<uint64_t>(), (uint32_t)n_tiles, ws, st));
    TileStats stats{};
    DI_CUDA(cudaMemcpyAsync(&stats, d_stats.p, sizeof stats, cudaMemcpyDeviceToHost, st));
    DI_CUDA(cudaStreamSynchronize(st));
    if (stats.bad_docid) return set_error(DI_ERR_RANGE, "term id >= n_terms (%u) in the collection", V);
    ix->n_postings = stats.n_visible;
    ix->max_docid_plus1 = ix->doc_lo + (uint32_t)n_docs;
    if (stats.n_visible == 0) return DI_OK;
    ix->n_tiles = (uint32_t)n_tiles;
    return finish_tiled(ix, ka.as<uint64_t>(), kb.as<uint64_t>(), ws.cur(), n_post, d_stats.as<TileStats>(), true, st);
}

static int new_index(uint32_t n_terms, uint32_t doc_lo, uint32_t doc_hi, const di_index_params *params, di_index **out)
{
    if (!out) return set_error(DI_ERR_ARG, "out is NULL");
    *out = nullptr;
    DI_TRY(ensure_device());
    if (doc_hi <= doc_lo) return set_error(DI_ERR_ARG, "empty doc range [%u, %u)", doc_lo, doc_hi);
    uint32_t tile_docs = params && params->tile_docs ? params->tile_docs : 16384u;
    if (tile_docs < 256 || tile_docs > kMaxTileDocs || (tile_docs & (tile_docs - 1)))
        return set_error(DI_ERR_ARG, "tile_docs must be a power of two in [256, %u], got %u", kMaxTileDocs, tile_docs);
    di_index *ix = new (std::nothrow) di_index();
    if (!ix) return set_error(DI_ERR_NOMEM, "host allocation failed");
    cudaGetDevice(&ix->device);
    ix->n_terms = n_terms;
    ix->doc_lo = doc_lo;
    ix->doc_hi = doc_hi;
    ix->tile_docs = tile_docs;
    while ((1u << ix->tile_shift) < tile_docs) ++ix->tile_shift;
    ix->dense_ratio = params && params->dense_ratio ? params->dense_ratio : 8u;
    ix->cand_slack = params ? params->cand_slack : 0u;
    ix->flags = params ? params->flags : 0u;
    if (getenv("DI_B200_NO_SEEDS")) ix->flags |= DI_INDEX_NO_SEEDS;   // tuning switches only
    if (getenv("DI_B200_PER_TILE")) ix->flags |= DI_INDEX_PER_TILE;
    if (getenv("DI_B200_NO_BANK_SORT")) ix->flags |= DI_INDEX_NO_BANK_SORT;
    cudaDeviceGetAttribute(&ix->smem_opt_in, cudaDevAttrMaxSharedMemoryPerBlockOptin, ix->device);
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ix;
        return set_error(DI_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    *out = ix;
    return DI_OK;
}

extern "C" int di_index_create_csr_dev(const uint64_t *d_term_offsets, const uint32_t *d_docids, const uint8_t *d_impacts,
                                       uint32_t n_terms, uint64_t n_postings, uint32_t doc_lo, uint32_t doc_hi,
                                       const di_index_params *params, di_index_t **out)
{
    di_index *ix = nullptr;
    DI_TRY(new_index(n_terms, doc_lo, doc_hi, params, &ix));
    int rc = build_tiled(ix, d_term_offsets, d_docids, d_impacts, n_postings, ix->stream);
    if (rc != DI_OK) {
        delete ix;
        *out = nullptr;
        return rc;
    }
    *out = ix;
    return DI_OK;
}

extern "C" int di_index_create_docmajor_dev(const uint32_t *d_term_ids, const uint8_t *d_impacts, const uint64_t *d_doc_offsets,
                                            uint64_t n_docs, uint32_t n_terms, uint64_t n_postings, uint32_t doc_lo,
                                            const di_index_params *params, di_index_t **out)
{
    if (n_docs == 0 || (uint64_t)doc_lo + n_docs > 0xFFFFFFFFull) return set_error(DI_ERR_ARG, "bad document range");
    di_index *ix = nullptr;
    DI_TRY(new_index(n_terms, doc_lo, doc_lo + (uint32_t)n_docs, params, &ix));
    int rc = build_tiled_docmajor(ix, d_term_ids, d_impacts, d_doc_offsets, n_docs, n_postings, ix->stream);
    if (rc != DI_OK) {
        delete ix;
        *out = nullptr;
        return rc;
    }
    *out = ix;
    return DI_OK;
}

extern "C" int di_index_create_csr(const uint64_t *term_offsets, const uint32_t *docids, const uint8_t *impacts,
                                   uint32_t n_terms, uint32_t doc_lo, uint32_t doc_hi, const di_index_params *params,
                                   di_index_t **out)
{
    if (!term_offsets || !out) return set_error(DI_ERR_ARG, "NULL argument");
    DI_TRY(ensure_device());
    const uint64_t P = term_offsets[n_terms];
    DevBuf d_to, d_d, d_v;
    DI_TRY(d_to.alloc(((size_t)n_terms + 1) * sizeof(uint64_t)));
    DI_TRY(d_d.alloc(P * sizeof(uint32_t)));
    DI_TRY(d_v.alloc(P));
    // the build runs on the index's own non-blocking stream, which is NOT ordered after the legacy stream:
    // finish the pageable copies before the first kernel can read them
    DI_CUDA(cudaMemcpy(d_to.p, term_offsets, ((size_t)n_terms + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    if (P) {
        DI_CUDA(cudaMemcpy(d_d.p, docids, P * sizeof(uint32_t), cudaMemcpyHostToDevice));
        DI_CUDA(cudaMemcpy(d_v.p, impacts, P, cudaMemcpyHostToDevice));
    }
    DI_CUDA(cudaDeviceSynchronize());
    return di_index_create_csr_dev(d_to.as<uint64_t>(), d_d.as<uint32_t>(), d_v.as<uint8_t>(), n_terms, P, doc_lo, doc_hi,
                                   params, out);
}

// Host -> device copy of a large PAGEABLE region (an mmap'ed index file): a single cudaMemcpy would fault the pages in and
// bounce them through the driver's staging buffer on one thread. Here the region is cut into chunks; host threads
// copy a chunk into one of a few PINNED staging buffers in parallel (that is what touches the file pages), the DMA of
// that chunk runs asynchronously while the threads already fill the next buffer. Returns after the last DMA finished.
static int upload_pinned(void *d_dst, const uint8_t *src, uint64_t bytes, cudaStream_t st)
{
    constexpr uint64_t kChunk = 32ull << 20;
    constexpr int kBufs = 4;
    if (bytes <= kChunk) {
        DI_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, st));
        DI_CUDA(cudaStreamSynchronize(st));
        return DI_OK;
    }
    uint8_t *stage[kBufs] = {};
    cudaEvent_t done[kBufs] = {};
    int rc = DI_OK;
    auto cleanup = [&] {
        for (int b = 0; b < kBufs; ++b) {
            if (done[b]) cudaEventDestroy(done[b]);
            if (stage[b]) cudaFreeHost(stage[b]);
        }
    };
    for (int b = 0; b < kBufs && rc == DI_OK; ++b) {
        if (cudaHostAlloc((void **)&stage[b], kChunk, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess)
            rc = set_error(DI_ERR_NOMEM, "pinned staging buffers for the index upload: %s", cudaGetErrorString(cudaGetLastError()));
    }
    const unsigned n_threads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    uint64_t chunk_i = 0;
    for (uint64_t off = 0; off < bytes && rc == DI_OK; off += kChunk, ++chunk_i) {
        const int b = (int)(chunk_i % kBufs);
        const uint64_t n = std::min(kChunk, bytes - off);
        if (chunk_i >= kBufs && cudaEventSynchronize(done[b]) != cudaSuccess) { rc = set_error(DI_ERR_CUDA, "index upload failed"); break; }
        {
            std::vector<std::thread> pool;
            const uint64_t per = (n + n_threads - 1) / n_threads;
            for (unsigned t = 0; t < n_threads; ++t) {
                const uint64_t lo = std::min<uint64_t>(n, t * per), hi = std::min<uint64_t>(n, lo + per);
                if (lo < hi) pool.emplace_back([=] { memcpy(stage[b] + lo, src + off + lo, hi - lo); });
            }
            for (std::thread &th : pool) th.join();
        }
        if (cudaMemcpyAsync((uint8_t *)d_dst + off, stage[b], n, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaEventRecord(done[b], st) != cudaSuccess)
            rc = set_error(DI_ERR_CUDA, "index upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == DI_OK) rc = set_error(DI_ERR_CUDA, "index upload failed");
    cleanup();
    return rc;
}

extern "C" int di_index_create_files(const uint8_t *dat, uint64_t dat_bytes, const uint64_t *idx_pairs, uint32_t n_terms,
                                     uint32_t doc_lo, uint32_t doc_hi, const di_index_params *params, di_index_t **out)
{
    if ((!dat && dat_bytes) || (!idx_pairs && n_terms) || !out) return set_error(DI_ERR_ARG, "NULL argument");
    DI_TRY(ensure_device());
    // inverted_index.py:47 — records are read while tell() < end: ceil((end - start) / 5) of them
    std::vector<uint64_t> offs((size_t)n_terms + 1, 0), starts((size_t)n_terms ? n_terms : 1, 0);
    for (uint32_t t = 0; t < n_terms; ++t) {
        const uint64_t s = idx_pairs[2 * (size_t)t], e = idx_pairs[2 * (size_t)t + 1];
        const uint64_t cnt = e > s ? (e - s + 4) / 5 : 0;
        if (cnt && s + 5 * cnt > dat_bytes)
            return set_error(DI_ERR_FORMAT, "term %u: records [%llu, %llu) run past the end of the .dat image (%llu bytes)", t,
                             (unsigned long long)s, (unsigned long long)(s + 5 * cnt), (unsigned long long)dat_bytes);
        offs[t + 1] = offs[t] + cnt;
        starts[t] = s;
    }
    const uint64_t P = offs[n_terms];
    DevBuf d_dat, d_to, d_st, d_d, d_v;
    DI_TRY(d_dat.alloc(dat_bytes));
    DI_TRY(d_to.alloc(((size_t)n_terms + 1) * sizeof(uint64_t)));
    DI_TRY(d_st.alloc((size_t)(n_terms ? n_terms : 1) * sizeof(uint64_t)));
    DI_TRY(d_d.alloc(P * sizeof(uint32_t)));
    DI_TRY(d_v.alloc(P));
    if (dat_bytes) DI_TRY(upload_pinned(d_dat.p, dat, dat_bytes, nullptr));
    DI_CUDA(cudaMemcpy(d_to.p, offs.data(), ((size_t)n_terms + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    if (n_terms) DI_CUDA(cudaMemcpy(d_st.p, starts.data(), (size_t)n_terms * sizeof(uint64_t), cudaMemcpyHostToDevice));
    if (P) {
        decode_dat_kernel<<<grid_for(P, 256), 256>>>(d_dat.as<uint8_t>(), d_to.as<uint64_t>(), d_st.as<uint64_t>(), n_terms, P,
                                                     d_d.as<uint32_t>(), d_v.as<uint8_t>());
        DI_KERNEL_CHECK();
        DI_CUDA(cudaDeviceSynchronize());
    }
    d_dat.release();
    return di_index_create_csr_dev(d_to.as<uint64_t>(), d_d.as<uint32_t>(), d_v.as<uint8_t>(), n_terms, P, doc_lo, doc_hi,
                                   params, out);
}

extern "C" void di_index_destroy(di_index_t *index)
{
    if (!index) return;
    cudaSetDevice(index->device);
    cudaDeviceSynchronize();
    delete index;
}

extern "C" int di_index_get_info(const di_index_t *ix, di_index_info *info)
{
    if (!ix || !info) return set_error(DI_ERR_ARG, "NULL argument");
    info->n_postings = ix->n_postings;
    info->payload_bytes = ix->payload_bytes;
    info->table_bytes = ix->table_bytes;
    info->n_dense_segments = ix->n_dense_segments;
    info->n_sparse_segments = ix->n_sparse_segments;
    info->n_dense_postings = ix->n_dense_postings;
    info->n_terms = ix->n_terms;
    info->doc_lo = ix->doc_lo;
    info->doc_hi = ix->doc_hi;
    info->n_tiles = ix->n_tiles;
    info->tile_docs = ix->tile_docs;
    info->max_docid_plus1 = ix->max_docid_plus1;
    return DI_OK;
}

extern "C" int di_index_export_seed_hist_dev(const di_index_t *ix, uint32_t *d_hist, void *stream)
{
    if (!ix || !d_hist) return set_error(DI_ERR_ARG, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    DI_CUDA(cudaSetDevice(ix->device));
    DI_CUDA(cudaMemsetAsync(d_hist, 0, (size_t)ix->n_terms * 256 * 4, st));
    const uint64_t n_segs = (uint64_t)ix->n_tiles * ix->n_terms;
    if (n_segs) {
        seed_export_kernel<<<grid_for(n_segs, 8, 148 * 32), 256, 0, st>>>(ix->d_desc, ix->d_payload, n_segs, ix->n_terms,
                                                                         ix->tile_docs, d_hist);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

extern "C" int di_index_import_seed_hist_dev(di_index_t *ix, const uint32_t *d_hist, void *stream)
{
    if (!ix || !d_hist) return set_error(DI_ERR_ARG, "NULL argument");
    if (ix->has_dup_postings) return DI_OK;  // k postings are not k documents here: such an index never seeds
    cudaStream_t st = (cudaStream_t)stream;
    DI_CUDA(cudaSetDevice(ix->device));
    DI_CUDA(cudaStreamSynchronize(ix->stream));  // no search of this index may still read the old tables
    if (ix->d_seed_cum) cudaFree(ix->d_seed_cum);
    if (ix->d_seed_slot) cudaFree(ix->d_seed_slot);
    ix->d_seed_cum = ix->d_seed_slot = nullptr;
    const uint32_t V = ix->n_terms;
    if (V == 0 || (ix->flags & DI_INDEX_NO_SEEDS)) return DI_OK;
    DevBuf d_cnt;
    DI_TRY(d_cnt.alloc(4));
    DI_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4, st));
    DI_CUDA(cudaMalloc(&ix->d_seed_slot, (size_t)V * 4));
    DI_CUDA(cudaMalloc(&ix->d_seed_cum, (size_t)V * 256 * 4));
    seed_import_kernel<<<(unsigned)(((uint64_t)V * 32 + 255) / 256), 256, 0, st>>>(d_hist, V, ix->d_seed_slot, d_cnt.as<uint32_t>(),
                                                                                 ix->d_seed_cum);
    DI_KERNEL_CHECK();
    uint32_t n_slots = 0;
    DI_CUDA(cudaMemcpyAsync(&n_slots, d_cnt.p, 4, cudaMemcpyDeviceToHost, st));
    DI_CUDA(cudaStreamSynchronize(st));
    if (n_slots) {
        seed_cum_kernel<<<(unsigned)(((uint64_t)n_slots * 32 + 255) / 256), 256, 0, st>>>(ix->d_seed_cum, n_slots);
        DI_KERNEL_CHECK();
    }
    DI_CUDA(cudaStreamSynchronize(st));
    ix->global_seeds = true;
    return DI_OK;
}

extern "C" int di_index_set_sorted_prefix(di_index_t *ix, uint32_t sorted_prefix)
{
    if (!ix) return set_error(DI_ERR_ARG, "index is NULL");
    ix->sorted_prefix = sorted_prefix;
    return DI_OK;
}

extern "C" int di_index_term_df(const di_index_t *ix, const uint32_t *term_ids, uint64_t n, uint64_t *df_out)
{
    if (!ix || (!term_ids && n) || (!df_out && n)) return set_error(DI_ERR_ARG, "NULL argument");
    std::vector<unsigned long long> df((size_t)ix->n_terms, 0);
    if (ix->d_df && ix->n_terms) {
        DI_CUDA(cudaSetDevice(ix->device));
        DI_CUDA(cudaMemcpy(df.data(), ix->d_df, (size_t)ix->n_terms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    for (uint64_t i = 0; i < n; ++i) df_out[i] = term_ids[i] < ix->n_terms ? df[term_ids[i]] : 0;
    return DI_OK;
}

// ============================================================================ search
static uint32_t pow2_ceil(uint32_t x)
{
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// finalize_topk_kernel with a shared-memory key buffer sized for the longest list it can meet
// (max_n), between 32 KB and 128 KB; longer lists take the kernel's global-memory path.
static int launch_finalize(uint64_t *cand, const uint32_t *cnt, uint32_t cap, uint32_t top_k,
                           uint32_t max_n, uint32_t n_queries, uint64_t *out_keys, uint32_t *out_counts, cudaStream_t st,
                           uint32_t sort_prefix = 0)
{
    // function attributes belong to a device: remember per device (not per thread) that the opt-in is done
    static std::atomic<uint64_t> attr_done{0};
    int dev = 0;
    DI_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
        DI_CUDA(cudaFuncSetAttribute(finalize_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        attr_done.fetch_or(bit, std::memory_order_release);
    }
    uint32_t smem_keys = kSortSmemKeys;
    while (smem_keys < max_n && smem_keys < 16384) smem_keys <<= 1;
    finalize_topk_kernel<<<n_queries, kScoreThreads, (size_t)smem_keys * 8, st>>>(cand, cnt, cap, top_k, smem_keys, sort_prefix,
                                                                                 out_keys, out_counts);
    DI_KERNEL_CHECK();
    return DI_OK;
}

static int ensure(DevBuf &b, size_t bytes)
{
    if (b.bytes >= bytes && b.p) return DI_OK;
    return b.alloc(bytes + bytes / 8);
}

extern "C" int di_search_dev(di_index_t *ix, const uint32_t *d_q_terms, const uint64_t *d_q_offsets, uint32_t n_queries,
                             uint32_t max_query_len, uint32_t top_k, const uint64_t *d_theta_init, uint64_t *d_out_keys,
                             uint32_t *d_out_counts, void *stream)
{
    if (!ix) return set_error(DI_ERR_ARG, "index is NULL");
    if (top_k == 0 || top_k > 65536) return set_error(DI_ERR_ARG, "top_k must be in [1, 65536], got %u", top_k);
    if (max_query_len > 65535) return set_error(DI_ERR_ARG, "queries longer than 65535 terms are not supported");
    cudaStream_t st = (cudaStream_t)stream;
    DI_CUDA(cudaSetDevice(ix->device));
    // timings accumulate over calls until di_get_timings() reads (and clears) them
    if (n_queries == 0) return DI_OK;

    // 257 * 255 = 65535 still fits a u16 accumulator — as long as a term adds to a document at most once
    const bool acc32 = max_query_len > 257 || ix->has_dup_postings;
    const size_t acc_bytes = score_smem_bytes(ix->tile_docs, acc32);  // dynamic shared memory per CTA
    if ((int)acc_bytes + 4096 > ix->smem_opt_in)
        return set_error(DI_ERR_ARG, "tile of %u docs needs %zu B of shared memory for %d-bit accumulators (limit %d); "
                         "rebuild the index with smaller tiles for queries of %u terms",
                         ix->tile_docs, acc_bytes, acc32 ? 32 : 16, ix->smem_opt_in, max_query_len);
    uint32_t c0 = ix->cand_slack ? ix->cand_slack : std::max(2u * top_k, 256u);
    c0 = std::max(c0, top_k);
    const uint32_t cap = std::max(c0 + ix->tile_docs, pow2_ceil(top_k));

    // four instantiations: 16- / 32-bit accumulators x with / without the tile bounds (an index without the bound
    // table runs a kernel that carries no trace of it)
    using PersistentFn = void (*)(SearchArgs, unsigned long long *);
    using TileFn = void (*)(SearchArgs, uint32_t);
    static const PersistentFn kPersistent[4] = {score_persistent_kernel<false, false>, score_persistent_kernel<false, true>,
                                                score_persistent_kernel<true, false>, score_persistent_kernel<true, true>};
    static const TileFn kTile[4] = {score_tile_kernel<false, false>, score_tile_kernel<false, true>,
                                    score_tile_kernel<true, false>, score_tile_kernel<true, true>};
    const int variant = (acc32 ? 2 : 0) | (ix->d_seg_max ? 1 : 0);
    if (!(ix->attr_set & (1u << variant))) {
        DI_CUDA(cudaFuncSetAttribute((const void *)kTile[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_bytes));
        DI_CUDA(cudaFuncSetAttribute((const void *)kPersistent[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_bytes));
        if (!acc32) {
            DI_CUDA(cudaFuncSetAttribute((const void *)kTile[variant], cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
            DI_CUDA(cudaFuncSetAttribute((const void *)kPersistent[variant], cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        }
        ix->attr_set |= 1u << variant;
    }
    const bool per_tile_launches = (ix->flags & DI_INDEX_PER_TILE) != 0;
    int ctas_per_sm = 0, n_sms = 0;
    DI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kPersistent[variant], kScoreThreads, acc_bytes));
    DI_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, ix->device));
    const int resident_ctas = std::max(1, ctas_per_sm) * std::max(1, n_sms);

    // candidate workspace: at most ~6 GB, queries are processed in batches that fit
    const uint64_t per_query = (uint64_t)cap * 8;
    uint32_t batch = (uint32_t)std::min<uint64_t>(n_queries, std::max<uint64_t>(1, (6ull << 30) / per_query));
    if ((n_queries + batch - 1) / batch > (uint32_t)di_index::kMaxBatches)
        batch = (n_queries + di_index::kMaxBatches - 1) / di_index::kMaxBatches;
    // A query is one chain of tiles; a batch smaller than the GPU (fewer queries than resident CTAs) is
    // widened by cutting the tile range into `lanes` independent sub-ranges per query, merged at the end.
    uint32_t lanes = 1;
    if (!per_tile_launches && ix->n_tiles > 1 && batch < (uint32_t)resident_ctas) {
        lanes = std::min<uint32_t>(ix->n_tiles, ((uint32_t)resident_ctas + batch - 1) / batch);
        lanes = std::min<uint32_t>(lanes, std::max<uint32_t>(1, 16384 / top_k));  // the lane merge stays in shared memory
        lanes = (uint32_t)std::min<uint64_t>(lanes, std::max<uint64_t>(1, (6ull << 30) / (per_query * batch)));
    }
    const uint32_t tiles_per_lane = ix->n_tiles ? (ix->n_tiles + lanes - 1) / lanes : 0;
    if (tiles_per_lane) lanes = (ix->n_tiles + tiles_per_lane - 1) / tiles_per_lane;
    const size_t n_virtual = (size_t)batch * lanes;
    ix->last_lanes = lanes;
    ix->last_acc32 = acc32 ? 1u : 0u;
    DI_TRY(ensure(ix->ws_cand, n_virtual * per_query));
    DI_TRY(ensure(ix->ws_cnt, n_virtual * 4));
    DI_TRY(ensure(ix->ws_theta, n_virtual * 8));
    DI_TRY(ensure(ix->ws_order, (size_t)batch * (sizeof(QueryRec) + 1)));  // records + one bucket byte per query
    DI_TRY(ensure(ix->ws_done, 8 + n_virtual * 4));
    if (ix->d_seg_max && !ix->ws_skipped.p) {
        DI_TRY(ix->ws_skipped.alloc(8));
        DI_CUDA(cudaMemsetAsync(ix->ws_skipped.p, 0, 8, st));
    }
    if (lanes > 1) {
        DI_TRY(ensure(ix->ws_lane_keys, n_virtual * top_k * 8));
        DI_TRY(ensure(ix->ws_lane_counts, n_virtual * 4));
    }

    for (uint32_t q0 = 0; q0 < n_queries; q0 += batch) {
        const uint32_t nq = std::min(batch, n_queries - q0);
        const size_t nv = (size_t)nq * lanes;
        if (ix->n_batches == di_index::kMaxBatches) {  // nobody is reading them: start over
            ix->n_batches = 0;
            ix->score_launches = ix->other_launches = 0;
        }
        const int b = ix->n_batches++;
        for (int e = 0; e < 3; ++e)
            if (!ix->ev[b][e]) DI_CUDA(cudaEventCreate(&ix->ev[b][e]));
        SearchArgs a{};
        a.desc = ix->d_desc;
        a.payload = ix->d_payload;
        a.q_terms = d_q_terms;
        a.q_offsets = d_q_offsets + q0;
        a.cand = ix->ws_cand.as<uint64_t>();
        a.cnt = ix->ws_cnt.as<uint32_t>();
        a.theta = ix->ws_theta.as<uint64_t>();
        a.n_terms = ix->n_terms;
        a.tile_docs = ix->tile_docs;
        a.tile_shift = ix->tile_shift;
        a.doc_lo = ix->doc_lo;
        a.cap = cap;
        a.c0 = c0;
        a.k = top_k;
        a.recs = ix->ws_order.as<QueryRec>();
        a.n_queries = nq;
        a.n_tiles = ix->n_tiles;
        a.seg_max = ix->d_seg_max;
        a.n_skipped = ix->d_seg_max ? ix->ws_skipped.as<unsigned long long>() : nullptr;
        a.lanes = lanes;
        a.tiles_per_lane = tiles_per_lane;
#ifdef DI_PROFILE_PHASES
        static DevBuf prof_buf;  // diagnostic build: per-tile phase cycles, dumped to $DI_B200_PROF after the launch
        const char *prof_path = getenv("DI_B200_PROF");
        a.prof = nullptr;
        if (prof_path) {
            DI_TRY(ensure(prof_buf, (size_t)ix->n_tiles * 64 + 64));
            DI_CUDA(cudaMemsetAsync(prof_buf.p, 0, (size_t)ix->n_tiles * 64, st));
            a.prof = prof_buf.as<unsigned long long>();
        }
#endif
        DI_CUDA(cudaMemsetAsync(a.cnt, 0, nv * 4, st));
        if (d_theta_init) {  // caller-proven lower bounds, one copy per lane
            for (uint32_t l = 0; l < lanes; ++l)
                DI_CUDA(cudaMemcpyAsync(a.theta + (size_t)l * nq, d_theta_init + q0, (size_t)nq * 8,
                                        cudaMemcpyDeviceToDevice, st));
        } else {
            DI_CUDA(cudaMemsetAsync(a.theta, 0, nv * 8, st));
        }
        DI_CUDA(cudaEventRecord(ix->ev[b][0], st));
        if (ix->n_tiles && ix->d_seed_cum) {
            seed_theta_kernel<<<(nq + 127) / 128, 128, 0, st>>>(d_q_terms, a.q_offsets, nq, lanes, ix->d_seed_slot,
                                                                ix->d_seed_cum, ix->n_terms, top_k, a.theta);
            DI_KERNEL_CHECK();
            ++ix->other_launches;
        }
        if (ix->n_tiles) {
            uint8_t *bucket = reinterpret_cast<uint8_t *>(ix->ws_order.as<QueryRec>() + batch);
            query_bucket_kernel<<<(nq + 127) / 128, 128, 0, st>>>(d_q_terms, a.q_offsets, ix->d_df, ix->n_terms, nq, bucket);
            DI_KERNEL_CHECK();
            query_order_kernel<<<1, 1024, 0, st>>>(d_q_terms, a.q_offsets, bucket, nq, ix->ws_order.as<QueryRec>());
            DI_KERNEL_CHECK();
            ix->other_launches += 2;
        }
        if (per_tile_launches) {  // DI_B200_PER_TILE=1: one launch per tile (to profile a single tile)
            a.done = nullptr;
            for (uint32_t tile = 0; tile < ix->n_tiles; ++tile) {
                kTile[variant]<<<nq, kScoreThreads, acc_bytes, st>>>(a, tile);
                ++ix->score_launches;
            }
        } else if (ix->n_tiles) {  // one persistent launch over all (lane, query, tile) items of the batch
            a.done = ix->ws_done.as<uint32_t>() + 2;  // [0..1] hold the 64-bit work counter
            unsigned long long *counter = ix->ws_done.as<unsigned long long>();
            DI_CUDA(cudaMemsetAsync(ix->ws_done.p, 0, 8 + nv * 4, st));
            const uint64_t n_items = (uint64_t)((tiles_per_lane + kTilesPerItem - 1) / kTilesPerItem) * nv;
            const unsigned grid = (unsigned)std::min<uint64_t>(n_items, (uint64_t)resident_ctas);
            kPersistent[variant]<<<grid, kScoreThreads, acc_bytes, st>>>(a, counter);
            ++ix->score_launches;
        }
        DI_KERNEL_CHECK();
        DI_CUDA(cudaEventRecord(ix->ev[b][1], st));
#ifdef DI_PROFILE_PHASES
        if (a.prof) {  // overwrite: the file holds the last launch
            std::vector<unsigned long long> h((size_t)ix->n_tiles * 8);
            DI_CUDA(cudaMemcpyAsync(h.data(), a.prof, h.size() * 8, cudaMemcpyDeviceToHost, st));
            DI_CUDA(cudaStreamSynchronize(st));
            if (FILE *f = fopen(prof_path, "w")) {
                fprintf(f, "tile,lookup,dense,sparse,wait_presel,scan_emit,cut,unused,items\n");
                for (uint32_t t = 0; t < ix->n_tiles; ++t) {
                    fprintf(f, "%u", t);
                    for (int c = 0; c < 8; ++c) fprintf(f, ",%llu", h[(size_t)t * 8 + c]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
#endif
        uint64_t *out_keys = d_out_keys + (uint64_t)q0 * top_k;
        uint32_t *out_counts = d_out_counts + q0;
        if (lanes == 1) {
            DI_TRY(launch_finalize(a.cand, a.cnt, cap, top_k, /*max_n=*/c0, nq, out_keys, out_counts, st, ix->sorted_prefix));
            ++ix->other_launches;
        } else {  // per-lane top-k rows [lanes][nq][k], then the same merge the multi-GPU path uses
            DI_TRY(launch_finalize(a.cand, a.cnt, cap, top_k, /*max_n=*/c0, (uint32_t)nv,
                                   ix->ws_lane_keys.as<uint64_t>(), ix->ws_lane_counts.as<uint32_t>(), st));
            DI_TRY(di_merge_topk_dev(ix->ws_lane_keys.as<uint64_t>(), ix->ws_lane_counts.as<uint32_t>(), lanes, nq, top_k, top_k,
                                     out_keys, out_counts, nullptr, st));
            ix->other_launches += 3;
        }
        DI_CUDA(cudaEventRecord(ix->ev[b][2], st));
    }
    return DI_OK;
}

extern "C" int di_unpack_keys_dev(const uint64_t *d_keys, uint64_t n, uint32_t *d_docids, int32_t *d_scores, void *stream)
{
    if (n == 0) return DI_OK;
    unpack_keys_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_keys, n, d_docids, d_scores);
    DI_KERNEL_CHECK();
    return DI_OK;
}

extern "C" int di_search(di_index_t *ix, const uint32_t *q_terms, const uint64_t *q_offsets, uint32_t n_queries,
                         uint32_t top_k, uint32_t *out_docids, int32_t *out_scores, uint32_t *out_counts)
{
    if (!ix || !q_offsets) return set_error(DI_ERR_ARG, "NULL argument");
    if (n_queries == 0) return DI_OK;
    if (top_k == 0 || top_k > 65536) return set_error(DI_ERR_ARG, "top_k must be in [1, 65536], got %u", top_k);
    uint64_t max_len = 0;
    for (uint32_t q = 0; q < n_queries; ++q) {
        if (q_offsets[q + 1] < q_offsets[q]) return set_error(DI_ERR_ARG, "q_offsets is not non-decreasing at %u", q);
        max_len = std::max(max_len, q_offsets[q + 1] - q_offsets[q]);
    }
    if (max_len > 65535) return set_error(DI_ERR_ARG, "queries longer than 65535 terms are not supported");
    const uint64_t base = q_offsets[0], n_terms_total = q_offsets[n_queries] - base;
    DI_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const uint64_t n_out = (uint64_t)n_queries * top_k;
    DI_TRY(ensure(ix->st_qterms, std::max<uint64_t>(n_terms_total, 1) * 4));
    DI_TRY(ensure(ix->st_qoffs, ((size_t)n_queries + 1) * 8));
    DI_TRY(ensure(ix->st_keys, n_out * 8));
    DI_TRY(ensure(ix->st_counts, (size_t)n_queries * 4));
    DI_TRY(ensure(ix->st_docids, n_out * 4));
    DI_TRY(ensure(ix->st_scores, n_out * 4));
    if (n_terms_total)
        DI_CUDA(cudaMemcpyAsync(ix->st_qterms.p, q_terms + base, n_terms_total * 4, cudaMemcpyHostToDevice, st));
    if (base == 0) {
        DI_CUDA(cudaMemcpyAsync(ix->st_qoffs.p, q_offsets, ((size_t)n_queries + 1) * 8, cudaMemcpyHostToDevice, st));
    } else {
        std::vector<uint64_t> rel((size_t)n_queries + 1);
        for (uint32_t q = 0; q <= n_queries; ++q) rel[q] = q_offsets[q] - base;
        DI_CUDA(cudaMemcpyAsync(ix->st_qoffs.p, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st));
        DI_CUDA(cudaStreamSynchronize(st));
    }
    if (ix->n_tiles == 0) {  // empty shard: nothing can match
        memset(out_counts, 0, (size_t)n_queries * 4);
        ix->n_batches = 0;
        DI_CUDA(cudaStreamSynchronize(st));
        return DI_OK;
    }
    const uint32_t row_order = ix->sorted_prefix;  // host rows are always fully sorted
    ix->sorted_prefix = 0;
    const int rc = di_search_dev(ix, ix->st_qterms.as<uint32_t>(), ix->st_qoffs.as<uint64_t>(), n_queries, (uint32_t)max_len, top_k,
                                 nullptr, ix->st_keys.as<uint64_t>(), ix->st_counts.as<uint32_t>(), st);
    ix->sorted_prefix = row_order;
    DI_TRY(rc);
    DI_TRY(di_unpack_keys_dev(ix->st_keys.as<uint64_t>(), n_out, ix->st_docids.as<uint32_t>(), ix->st_scores.as<int32_t>(), st));
    ++ix->other_launches;
    DI_CUDA(cudaMemcpyAsync(out_docids, ix->st_docids.p, n_out * 4, cudaMemcpyDeviceToHost, st));
    DI_CUDA(cudaMemcpyAsync(out_scores, ix->st_scores.p, n_out * 4, cudaMemcpyDeviceToHost, st));
    DI_CUDA(cudaMemcpyAsync(out_counts, ix->st_counts.p, (size_t)n_queries * 4, cudaMemcpyDeviceToHost, st));
    DI_CUDA(cudaStreamSynchronize(st));
    return DI_OK;
}

extern "C" int di_merge_topk_dev(const uint64_t *d_keys_in, const uint32_t *d_counts_in, uint32_t n_shards,
                                 uint32_t n_queries, uint32_t k_in, uint32_t top_k, uint64_t *d_keys_out,
                                 uint32_t *d_counts_out, uint32_t *d_incomplete, void *stream)
{
    if (n_queries == 0 || n_shards == 0) return DI_OK;
    if (top_k == 0 || top_k > 65536) return set_error(DI_ERR_ARG, "top_k must be in [1, 65536], got %u", top_k);
    if (k_in == 0 || k_in > 65536) return set_error(DI_ERR_ARG, "k_in must be in [1, 65536], got %u", k_in);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t cap = pow2_ceil(std::max(n_shards * k_in, top_k));
    // scratch comes from the device's stream-ordered pool: it belongs to this call's device and stream, so
    // merges on several devices or streams of one thread cannot share (and race on) it; after the first call
    // the pool serves it without a driver allocation and the call stays asynchronous
    StreamBuf cand(st), cnt(st);
    DI_TRY(cand.alloc((size_t)n_queries * cap * 8));
    DI_TRY(cnt.alloc((size_t)n_queries * 4));
    merge_gather_kernel<<<n_queries, 256, 0, st>>>(d_keys_in, d_counts_in, n_shards, n_queries, k_in, cand.as<uint64_t>(),
                                                  cnt.as<uint32_t>(), cap);
    DI_KERNEL_CHECK();
    DI_TRY(launch_finalize(cand.as<uint64_t>(), cnt.as<uint32_t>(), cap, top_k, /*max_n=*/n_shards * k_in, n_queries,
                           d_keys_out, d_counts_out, st));
    if (d_incomplete) {
        merge_check_kernel<<<grid_for(n_queries, 256), 256, 0, st>>>(d_keys_in, d_counts_in, n_shards, n_queries, k_in, top_k,
                                                                    d_keys_out, d_counts_out, d_incomplete);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

extern "C" int di_merge_pull_dev(const uint64_t *const *d_rows, const uint32_t *const *d_counts, uint32_t n_shards,
                                 uint32_t q_first, uint32_t n_queries, uint32_t row_stride, uint32_t k_in, uint32_t top_k,
                                 uint64_t *d_keys_out, uint32_t *d_counts_out, uint32_t *d_n_second_pass, void *stream)
{
    if (n_queries == 0 || n_shards == 0) return DI_OK;
    if (!d_rows || !d_counts) return set_error(DI_ERR_ARG, "NULL shard pointer table");
    if (n_shards > kMaxPeerShards) return set_error(DI_ERR_ARG, "at most %u shards, got %u", kMaxPeerShards, n_shards);
    if (top_k == 0 || top_k > 65536) return set_error(DI_ERR_ARG, "top_k must be in [1, 65536], got %u", top_k);
    if (k_in == 0 || row_stride == 0) return set_error(DI_ERR_ARG, "k_in and row_stride must be positive");
    const uint32_t lim1 = std::min(k_in, row_stride);
    const uint32_t smem1 = pow2_ceil((uint32_t)std::min<uint64_t>((uint64_t)n_shards * lim1 + pow2_ceil(top_k), 1u << 20));
    // the second pass of a query holds every shard's full row in shared memory
    const uint32_t smem2 = pow2_ceil((uint32_t)std::min<uint64_t>((uint64_t)n_shards * std::min(row_stride, top_k) + pow2_ceil(top_k), 1u << 20));
    int dev = 0, smem_max = 0;
    DI_CUDA(cudaGetDevice(&dev));
    DI_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (row_stride > top_k || (uint64_t)smem2 * 8 + 4096 > (uint64_t)smem_max)
        return set_error(DI_ERR_ARG, "%u shards x rows of %u keys do not fit the merge kernel's shared memory; "
                         "gather the rows and use di_merge_topk_dev", n_shards, row_stride);
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
        DI_CUDA(cudaFuncSetAttribute(merge_pull_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 4096));
        DI_CUDA(cudaFuncSetAttribute(merge_pull_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max - 4096));
        attr_done.fetch_or(bit, std::memory_order_release);
    }
    cudaStream_t st = (cudaStream_t)stream;
    StreamBuf redo(st);
    DI_TRY(redo.alloc((size_t)n_queries * 4));
    merge_pull_kernel<false><<<n_queries, kMergeThreads, (size_t)smem1 * 8, st>>>(
        d_rows, d_counts, n_shards, q_first, row_stride, k_in, top_k, smem1, d_keys_out, d_counts_out, redo.as<uint32_t>(),
        d_n_second_pass);
    DI_KERNEL_CHECK();
    if (lim1 < row_stride) {
        merge_pull_kernel<true><<<n_queries, kMergeThreads, (size_t)smem2 * 8, st>>>(
            d_rows, d_counts, n_shards, q_first, row_stride, k_in, top_k, smem2, d_keys_out, d_counts_out, redo.as<uint32_t>(),
            d_n_second_pass);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

// ---- peer-visible memory + cross-GPU barrier (one process per GPU; CUDA IPC over NVLink / NVSwitch)
extern "C" int di_shared_alloc(uint64_t bytes, void **d_ptr, uint8_t handle_out[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    if (!d_ptr || !handle_out) return set_error(DI_ERR_ARG, "NULL argument");
    DI_TRY(ensure_device());
    void *p = nullptr;
    DI_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    DI_CUDA(cudaMemset(p, 0, bytes ? bytes : 16));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return set_error(DI_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    *d_ptr = p;
    return DI_OK;
}

extern "C" int di_shared_open(const uint8_t handle[64], void **d_ptr)
{
    if (!d_ptr || !handle) return set_error(DI_ERR_ARG, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    DI_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DI_OK;
}

extern "C" int di_shared_close(void *d_ptr)
{
    if (d_ptr) DI_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return DI_OK;
}

extern "C" int di_shared_free(void *d_ptr)
{
    if (d_ptr) DI_CUDA(cudaFree(d_ptr));
    return DI_OK;
}

// page-locked host memory for result buffers: a device-to-host copy into it runs at full PCIe rate instead of being
// bounced through the driver's staging buffer (and the pages are never faulted in again)
extern "C" int di_host_alloc(uint64_t bytes, void **ptr)
{
    if (!ptr) return set_error(DI_ERR_ARG, "NULL argument");
    DI_TRY(ensure_device());
    DI_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return DI_OK;
}

extern "C" int di_host_free(void *ptr)
{
    if (ptr) DI_CUDA(cudaFreeHost(ptr));
    return DI_OK;
}

extern "C" int di_peer_barrier_dev(uint32_t *const *d_flags, uint32_t n_ranks, uint32_t my_rank, uint32_t epoch, void *stream)
{
    if (!d_flags || n_ranks == 0 || n_ranks > kMaxPeerShards || my_rank >= n_ranks)
        return set_error(DI_ERR_ARG, "bad barrier arguments (%u ranks, rank %u)", n_ranks, my_rank);
    peer_barrier_kernel<<<1, kMaxPeerShards, 0, (cudaStream_t)stream>>>(d_flags, n_ranks, my_rank, epoch);
    DI_KERNEL_CHECK();
    return DI_OK;
}

extern "C" int di_get_timings(di_index_t *ix, di_timings *out)
{
    if (!ix || !out) return set_error(DI_ERR_ARG, "NULL argument");
    memset(out, 0, sizeof *out);
    DI_CUDA(cudaSetDevice(ix->device));
    for (int b = 0; b < ix->n_batches; ++b) {
        float s = 0, f = 0;
        DI_CUDA(cudaEventSynchronize(ix->ev[b][2]));
        DI_CUDA(cudaEventElapsedTime(&s, ix->ev[b][0], ix->ev[b][1]));
        DI_CUDA(cudaEventElapsedTime(&f, ix->ev[b][1], ix->ev[b][2]));
        out->score_ms += s;
        out->finalize_ms += f;
    }
    if (ix->n_batches) {
        float t = 0;
        DI_CUDA(cudaEventElapsedTime(&t, ix->ev[0][0], ix->ev[ix->n_batches - 1][2]));
        out->total_ms = t;
    }
    out->score_launches = ix->score_launches;
    out->other_launches = ix->other_launches;
    out->lanes = ix->last_lanes;
    out->acc32 = ix->last_acc32;
    if (ix->ws_skipped.p) {   // tiles skipped since the previous read (the events above have synchronised the stream's work)
        unsigned long long n = 0;
        DI_CUDA(cudaMemcpy(&n, ix->ws_skipped.p, 8, cudaMemcpyDeviceToHost));
        DI_CUDA(cudaMemset(ix->ws_skipped.p, 0, 8));
        out->tiles_skipped = n;
    }
    ix->n_batches = 0;
    ix->score_launches = ix->other_launches = 0;
    return DI_OK;
}
