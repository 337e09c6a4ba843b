// build.cuh — index build kernels: K1 quantization, K2 term->document inversion + byte-exact
// serialisation, and the re-layout of a shard into document tiles (DESIGN.md "HBM layout").
#pragma once

#include "common.cuh"
#include "scan_sort.cuh"

namespace di {

// ============================================================================ K1: quantize
// quantize.py:17-24 — max over all scores, seeded with 0.0. Doubles compare like their bit
// patterns when non-negative, so a 64-bit integer atomicMax on the bits of max(v, 0) is exact.
__global__ void max_f64_kernel(const double *__restrict__ x, int64_t n, unsigned long long *__restrict__ out_bits)
{
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        if (v > m) m = v;  // NaN never wins, as in Python's max(max_val, nan) with max_val first
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double t = __shfl_xor_sync(0xffffffffu, m, o);
        if (t > m) m = t;
    }
    if (lane_id() == 0 && m > 0.0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
}

// quantize.py:13-14 — int(value * scale): ONE float64 multiply (no FMA to contract with: the
// product is the only operation) and truncation toward zero, saturated to int32.
__global__ void quantize_f64_kernel(const double *__restrict__ x, int64_t n, double scale, int32_t *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double p = __dmul_rn(x[i], scale);
        out[i] = __double2int_rz(p);  // saturating; NaN -> 0
    }
}

// ============================================================================ K2: inversion
// key = [term:24 | 255-impact:8 | docid:32]; the input is doc-major, i.e. docid ascending, so
// a STABLE sort on the top 32 bits yields term asc, impact desc, docid asc == create.py:33-41.
constexpr int kInvTermShift = 40;
constexpr uint32_t kMaxTerms = 1u << 24;

__device__ __forceinline__ uint64_t upper_bound_u64(const uint64_t *a, uint64_t n, uint64_t v)
{
    // first index with a[idx] > v, over a[0..n)
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The sort's digit counts are taken here, while the key is in a register (the sort would otherwise read all keys once
// more just to count them): sort_hist = [n_passes][256] counters of radix_sort_u64's bits [sort_lo, sort_hi).
__global__ void __launch_bounds__(256) invert_keys_kernel(const uint32_t *__restrict__ term_ids, const uint8_t *__restrict__ impacts,
                                                          const uint64_t *__restrict__ doc_offsets, uint64_t n_docs, uint32_t n_terms,
                                                          uint64_t *__restrict__ keys, uint32_t *__restrict__ status,
                                                          uint32_t *__restrict__ sort_hist, int sort_lo, int sort_hi, int n_passes)
{
    __shared__ uint32_t s_counts[4][256];
    for (int i = threadIdx.x; i < n_passes * 256; i += blockDim.x) (&s_counts[0][0])[i] = 0;
    __syncthreads();
    // one warp per document (lists are ~100 postings): the docid comes for free and the accesses stay
    // coalesced, instead of a 23-step binary search over doc_offsets per posting
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t doc = warp; doc < n_docs; doc += n_warps) {
        const uint64_t lo = doc_offsets[doc], hi = doc_offsets[doc + 1];
        for (uint64_t i = lo + lane_id(); i < hi; i += 32) {
            uint32_t t = term_ids[i];
            if (t >= n_terms) {  // not a term of the vocabulary: sorts behind every list and is left out; the caller is told
                t = n_terms;
                if (status) *status = 1u;
            }
            const uint64_t key = ((uint64_t)t << kInvTermShift) | ((uint64_t)(255u - impacts[i]) << 32) | (uint64_t)(uint32_t)doc;
            keys[i] = key;
            rs_count_key(s_counts, key, sort_lo, sort_hi, n_passes);
        }
    }
    __syncthreads();
    rs_flush_counts(s_counts, n_passes, sort_hist);
}

// `skip_if` (device word, may be nullptr): non-zero = the sort's last pass has already stored docids / impacts itself.
__global__ void invert_extract_kernel(const uint64_t *__restrict__ ka, const uint64_t *__restrict__ kb,
                                      const uint32_t *__restrict__ cur, uint64_t n_post, uint32_t n_terms,
                                      uint64_t *__restrict__ term_offsets, uint32_t *__restrict__ docids,
                                      uint8_t *__restrict__ impacts, const uint32_t *__restrict__ skip_if)
{
    if (skip_if && *skip_if) return;
    const uint64_t *__restrict__ keys = rs_result(ka, kb, cur);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        docids[i] = (uint32_t)k;
        impacts[i] = (uint8_t)(255u - ((uint32_t)(k >> 32) & 255u));
        const int64_t t = (int64_t)(k >> kInvTermShift);  // <= n_terms (n_terms = postings of unknown terms, left out)
        const int64_t tprev = i ? (int64_t)(keys[i - 1] >> kInvTermShift) : -1;
        for (int64_t tt = tprev + 1; tt <= t; ++tt) term_offsets[tt] = i;  // also covers empty terms
        if (i == n_post - 1)
            for (int64_t tt = t + 1; tt <= (int64_t)n_terms; ++tt) term_offsets[tt] = n_post;
    }
}

// term_offsets[t] = first position of a term >= t = suffix minimum of the first positions the sort's last pass noted
// (terms without postings have none), n_post behind the last one. One block; `run_if`: see invert_extract_kernel.
__global__ void __launch_bounds__(1024) invert_offsets_kernel(const unsigned long long *__restrict__ first, uint32_t n_terms,
                                                            uint64_t n_post, uint64_t *__restrict__ term_offsets,
                                                            const uint32_t *__restrict__ run_if)
{
    if (!*run_if) return;
    __shared__ unsigned long long s_min[1024];
    const uint32_t n = n_terms + 1, per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
    unsigned long long m = n_post;
    for (uint32_t t = hi; t > lo; --t) m = min(m, first[t - 1]);
    s_min[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0)  // suffix minimum over the chunks (1024 steps)
        for (int c = (int)blockDim.x - 2; c >= 0; --c) s_min[c] = min(s_min[c], s_min[c + 1]);
    __syncthreads();
    m = threadIdx.x + 1 < blockDim.x ? s_min[threadIdx.x + 1] : n_post;
    for (uint32_t t = hi; t > lo; --t) {
        m = min(m, first[t - 1]);
        term_offsets[t - 1] = m;
    }
}

// create.py:44-51 — 5-byte records and (start,end) byte offsets
__global__ void serialize_dat_kernel(const uint32_t *__restrict__ docids, const uint8_t *__restrict__ impacts,
                                     uint64_t n_post, uint8_t *__restrict__ dat)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t d = docids[i];
        uint8_t *p = dat + 5 * i;
        p[0] = (uint8_t)d; p[1] = (uint8_t)(d >> 8); p[2] = (uint8_t)(d >> 16); p[3] = (uint8_t)(d >> 24);
        p[4] = impacts[i];
    }
}

__global__ void serialize_idx_kernel(const uint64_t *__restrict__ term_offsets, uint32_t n_terms, uint64_t *__restrict__ idx)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += gridDim.x * blockDim.x) {
        idx[2 * (uint64_t)t] = 5 * term_offsets[t];
        idx[2 * (uint64_t)t + 1] = 5 * term_offsets[t + 1];
    }
}

// Asynchronous on `st`: scratch comes from the stream-ordered pool (no driver allocation after the first call) and
// nothing waits for the GPU. *d_status (optional) is set to 1 when a term id >= n_terms was met; those postings are
// left out and term_offsets[n_terms] is the number of postings kept.
inline int invert_dev(const uint32_t *d_term_ids, const uint8_t *d_impacts, const uint64_t *d_doc_offsets,
                      uint64_t n_docs, uint32_t n_terms, uint64_t n_post, uint64_t *d_term_offsets,
                      uint32_t *d_out_docids, uint8_t *d_out_impacts, uint32_t *d_status, cudaStream_t st)
{
    if (n_terms >= kMaxTerms) return set_error(DI_ERR_ARG, "n_terms %u exceeds 2^24 - 1", n_terms);
    if (n_docs >= (1ull << 32)) return set_error(DI_ERR_ARG, "n_docs exceeds 2^32-1");
    if (d_status) DI_CUDA(cudaMemsetAsync(d_status, 0, sizeof(uint32_t), st));
    if (n_post == 0) {
        DI_CUDA(cudaMemsetAsync(d_term_offsets, 0, ((size_t)n_terms + 1) * sizeof(uint64_t), st));
        return DI_OK;
    }
    StreamBuf ka(st), kb(st);
    RadixSortScratch ws(st);
    DI_TRY(ka.alloc(n_post * sizeof(uint64_t)));
    DI_TRY(kb.alloc(n_post * sizeof(uint64_t)));
    int term_bits = 1;
    while ((1ull << term_bits) < (uint64_t)n_terms + 1) ++term_bits;  // the value n_terms marks unknown terms
    const int sort_lo = 32, sort_hi = kInvTermShift + term_bits, n_passes = (sort_hi - sort_lo + 7) / 8;  // <= 4 passes (24 + 8 bits)
    DI_TRY(rs_prepare_counts(ws, n_passes, st));
    invert_keys_kernel<<<grid_for(n_post, 256, 148 * 8), 256, 0, st>>>(d_term_ids, d_impacts, d_doc_offsets, n_docs, n_terms,
                                                                       ka.as<uint64_t>(), d_status, ws.hist.as<uint32_t>(), sort_lo,
                                                                       sort_hi, n_passes);
    DI_KERNEL_CHECK();
    // the sort's last pass stores docids / impacts itself and notes where every term starts (RsEpilogue); the separate
    // pass over the sorted keys only runs when the device skipped that pass (every posting in one digit)
    StreamBuf first(st);
    DI_TRY(first.alloc(((size_t)n_terms + 1) * sizeof(unsigned long long)));
    DI_CUDA(cudaMemsetAsync(first.p, 0xFF, ((size_t)n_terms + 1) * sizeof(unsigned long long), st));
    RsEpilogue epi;
    epi.low = d_out_docids;
    epi.byte = d_out_impacts;
    epi.first = first.as<unsigned long long>();
    epi.field_shift = kInvTermShift;
#ifdef DI_INVERT_NO_EPILOGUE
    const RsEpilogue *use_epi = nullptr;
#else
    const RsEpilogue *use_epi = &epi;
#endif
    DI_TRY(radix_sort_u64(ka.as<uint64_t>(), kb.as<uint64_t>(), n_post, sort_lo, sort_hi, nullptr, 1, ws, st, /*precounted=*/true,
                          use_epi));
    const uint32_t *fused = use_epi && n_post > 1 ? ws.needed(n_passes - 1) : nullptr;
    invert_extract_kernel<<<grid_for(n_post, 256), 256, 0, st>>>(ka.as<uint64_t>(), kb.as<uint64_t>(), ws.cur(), n_post,
                                                                 n_terms, d_term_offsets, d_out_docids, d_out_impacts, fused);
    DI_KERNEL_CHECK();
    if (fused) {
        invert_offsets_kernel<<<1, 1024, 0, st>>>(first.as<unsigned long long>(), n_terms, n_post, d_term_offsets, fused);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

// ============================================================================ tiled shard layout
// tile key = [tile:16 | term:24 | parity:1 | local docid >> 1 : 15 | impact:8]; hidden postings carry a term
// field >= n_terms (~0 from term-major input: they sort behind everything; n_terms from doc-major input: they sort
// behind the visible postings of their tile) and are skipped by every consumer. Sorting puts a segment's even documents first, then its odd documents, each
// by ascending docid: the scorer adds a u8 impact into a u16 accumulator that shares a 32-bit
// shared-memory word with its neighbour, and knowing the parity per 16-byte unit makes the
// addend a single instruction (DESIGN.md "sparse segments").
constexpr int kTkTermShift = 24, kTkTileShift = 48, kTkLocalShift = 8;
constexpr uint32_t kDenseFlag = 0x80000000u;
constexpr int kDenseUnitShift = 4;  // a dense segment holds one byte per document: 16 documents per 16-byte unit
constexpr uint32_t kMaxTileDocs = 32768;

struct SegDesc {        // one per (tile, term)
    uint32_t off16;     // payload offset in 16-byte units
    uint32_t n_flag;    // dense: kDenseFlag | postings.  sparse: even units << 16 | total units (16 B units)
};

__host__ __device__ __forceinline__ uint32_t local_to_field(uint32_t local) { return ((local & 1u) << 15) | (local >> 1); }
__host__ __device__ __forceinline__ uint32_t field_to_local(uint32_t f) { return ((f & 0x7FFFu) << 1) | (f >> 15); }

// inverted_index.py:50-51 — the reader stops at the FIRST zero impact of a term's list
__global__ void first_zero_kernel(const uint64_t *__restrict__ term_offsets, uint32_t n_terms,
                                  const uint8_t *__restrict__ impacts, uint64_t n_post,
                                  unsigned long long *__restrict__ first_zero)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        if (impacts[i] == 0) {
            const uint64_t t = upper_bound_u64(term_offsets, (uint64_t)n_terms + 1, i) - 1;
            atomicMin(&first_zero[t], (unsigned long long)i);
        }
    }
}

__global__ void init_first_zero_kernel(const uint64_t *__restrict__ term_offsets, uint32_t n_terms,
                                       unsigned long long *__restrict__ first_zero)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += gridDim.x * blockDim.x)
        first_zero[t] = term_offsets[t + 1];
}

struct TileStats {
    unsigned long long n_visible;
    unsigned long long n_dense_segments, n_sparse_segments, n_dense_postings;
    unsigned int max_docid_plus1;
    unsigned int bad_docid;
    unsigned int n_dup_segments;  // segments listing one document twice (hand-made CSR): no threshold seeding then
};

// ---------------------------------------------------------------------------- threshold seeds
// A query's k-th best score is at least the k-th highest impact of any single one of its terms: k
// different documents hold that term with at least that impact, and the other terms only add.
// So a per-term table cum[v] = #visible postings with impact >= v gives every (query, k) a proven lower
// bound on its k-th best score before a single posting is scored: the first tiles of a search (and
// of every shard and lane of it) then start with a threshold instead of flooding their candidate
// lists. Exhaustive scoring (inverted_index.py:57-62) is unchanged: documents below a proven bound
// can never be in the top k. Terms with fewer than kSeedMinDf postings get no table (bound 0): they
// cannot flood anything.
constexpr uint32_t kSeedMinDf = 4096;
constexpr uint32_t kSeedChunk = 16384;   // postings per CTA of impact_hist_kernel
constexpr uint32_t kNoSeedSlot = 0xFFFFFFFFu;

__global__ void seed_slots_kernel(const uint64_t *__restrict__ term_offsets, uint32_t n_terms,
                                  uint32_t *__restrict__ slot_of_term, uint32_t *__restrict__ counter)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += gridDim.x * blockDim.x)
        slot_of_term[t] = term_offsets[t + 1] - term_offsets[t] >= kSeedMinDf ? atomicAdd(counter, 1u) : kNoSeedSlot;
}

// hist[slot][v] += #visible postings of the slot's term with impact v, for the chunk of the term-major
// posting arrays this CTA owns. Visible = what tile_keys_kernel keeps: before the term's first zero
// impact and inside the shard's docid range.
__global__ void __launch_bounds__(256) impact_hist_kernel(const uint64_t *__restrict__ term_offsets, uint32_t n_terms,
                                                          const uint32_t *__restrict__ docids,
                                                          const uint8_t *__restrict__ impacts, uint64_t n_post,
                                                          const unsigned long long *__restrict__ first_zero, uint32_t doc_lo,
                                                          uint32_t doc_hi, const uint32_t *__restrict__ slot_of_term,
                                                          uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_h[256];
    const uint64_t i0 = (uint64_t)blockIdx.x * kSeedChunk, i1 = min(i0 + (uint64_t)kSeedChunk, n_post);
    if (i0 >= i1) return;
    const uint64_t t_first = upper_bound_u64(term_offsets, (uint64_t)n_terms + 1, i0) - 1;
    const uint64_t t_last = upper_bound_u64(term_offsets, (uint64_t)n_terms + 1, i1 - 1) - 1;
    for (uint64_t t = t_first; t <= t_last; ++t) {  // uniform across the CTA
        const uint32_t slot = slot_of_term[t];
        if (slot == kNoSeedSlot) continue;
        const uint64_t lo = max(term_offsets[t], i0), hi = min(min(term_offsets[t + 1], i1), (uint64_t)first_zero[t]);
        s_h[threadIdx.x] = 0;
        __syncthreads();
        // Lists written by the inversion are sorted by impact (create.py:41), so the 32 postings of a warp mostly
        // share one value: count RUNS (one atomic per run) instead of 32 atomics on the same bin.
        const uint32_t lane = lane_id();
        for (uint64_t base = lo + (threadIdx.x - lane); base < hi; base += 256) {  // warp-uniform
            const uint64_t i = base + lane;
            uint32_t v = 0x100u + lane;  // not counted, and never equal to a neighbour
            if (i < hi) {
                const uint32_t d = docids[i];
                if (d >= doc_lo && d < doc_hi) v = impacts[i];
            }
            const uint32_t prev = __shfl_up_sync(0xffffffffu, v, 1);
            const bool head = lane == 0 || v != prev;
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            if (head && v < 0x100u) {
                const uint32_t next = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
                atomicAdd(&s_h[v], (next ? (uint32_t)__ffs(next) - 1u : 32u) - lane);
            }
        }
        __syncthreads();
        if (s_h[threadIdx.x]) atomicAdd(&hist[(size_t)slot * 256 + threadIdx.x], s_h[threadIdx.x]);
        __syncthreads();
    }
}

// in place: hist[slot][v] -> #postings with impact >= v
__global__ void seed_cum_kernel(uint32_t *__restrict__ hist, uint32_t n_slots)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (warp >= n_slots) return;
    uint32_t *h = hist + (size_t)warp * 256;
    uint32_t c[8], sum = 0;  // lane L owns bins 255-8L .. 248-8L, descending
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sum += h[255 - 8 * lane - j];
        c[j] = sum;
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += up;
    }
    const uint32_t before = incl - sum;
#pragma unroll
    for (int j = 0; j < 8; ++j) h[255 - 8 * lane - j] = before + c[j];
}

// theta[lane][q] = max(theta[lane][q], bound(q) << 32) with bound(q) = max over the query's terms of the largest
// v with cum[term][v] >= k (0 when no term has k postings). One thread per query.
__global__ void seed_theta_kernel(const uint32_t *__restrict__ q_terms, const uint64_t *__restrict__ q_offsets,
                                  uint32_t n_queries, uint32_t lanes, const uint32_t *__restrict__ slot_of_term,
                                  const uint32_t *__restrict__ cum, uint32_t n_terms, uint32_t k, uint64_t *__restrict__ theta)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    uint32_t best = 0;
    for (uint64_t j = q_offsets[q]; j < q_offsets[q + 1]; ++j) {
        const uint32_t t = q_terms[j];
        if (t >= n_terms) continue;
        const uint32_t slot = slot_of_term[t];
        if (slot == kNoSeedSlot) continue;
        const uint32_t *c = cum + (size_t)slot * 256;
        if (c[best + 1 > 255 ? 255 : best + 1] < k) continue;  // cannot improve on the current bound
        uint32_t lo = best, hi = 255;                          // invariant: cum[lo] >= k (or lo == best), cum[hi + 1] < k
        while (lo < hi) {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (c[mid] >= k) lo = mid; else hi = mid - 1;
        }
        best = lo;
    }
    if (best == 0) return;
    const uint64_t bound = (uint64_t)best << 32;  // (score, worst possible docid): below every key of that score
    for (uint32_t l = 0; l < lanes; ++l) {
        uint64_t *p = theta + (size_t)l * n_queries + q;
        if (*p < bound) *p = bound;
    }
}

// ---- seeds of a SHARDED collection ------------------------------------------------------------------------------------
// The seed of a query is a lower bound of the k-th best score over the documents the tables count. A shard that only knows
// its own postings can only bound its OWN k-th score; with the tables of all shards added up every shard starts from a
// bound of the GLOBAL k-th score and never emits what the cross-shard merge would throw away (measured at 8 shards:
// score kernel 4.06 -> 3.86 ms, finalize 0.23 -> 0.20 ms). export: this shard's impact histogram of EVERY term as a dense
// [n_terms][256] table (read back from the tiled payload; the caller sums the tables of all shards, e.g. with one
// all-reduce); import: new seed tables from such a sum.
__global__ void __launch_bounds__(256) seed_export_kernel(const SegDesc *__restrict__ desc, const uint8_t *__restrict__ payload,
                                                          uint64_t n_segs, uint32_t n_terms, uint32_t tile_docs,
                                                          uint32_t *__restrict__ hist /* [n_terms][256] */)
{
    __shared__ uint32_t s_h[8][256];
    const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
    const uint4 *payload4 = reinterpret_cast<const uint4 *>(payload);
    for (uint64_t seg = (uint64_t)blockIdx.x * 8 + w; seg < n_segs; seg += (uint64_t)gridDim.x * 8) {
        const SegDesc d = desc[seg];
        if (d.n_flag == 0) continue;
        uint32_t *h = hist + (size_t)(seg % n_terms) * 256;
        if (d.n_flag & kDenseFlag) {  // one byte per document, 0 = absent: count in shared memory, flush the bins in use
            for (int i = lane; i < 256; i += 32) s_h[w][i] = 0;
            __syncwarp();
            for (uint32_t u = lane; u < (tile_docs >> kDenseUnitShift); u += 32) {
                const uint4 v = payload4[(size_t)d.off16 + u];
                const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t b = (x[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                    if (b) atomicAdd(&s_h[w][b], 1u);
                }
            }
            __syncwarp();
            for (int i = lane; i < 256; i += 32)
                if (s_h[w][i]) atomicAdd(&h[i], s_h[w][i]);
            __syncwarp();
        } else {  // u32 postings, impact in bits 16..23; padding words carry impact 0
            const uint32_t total = d.n_flag & 0xFFFFu;
            for (uint32_t u = lane; u < total; u += 32) {
                const uint4 v = payload4[(size_t)d.off16 + u];
                const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t b = (x[i] >> 16) & 0xFFu;
                    if (b) atomicAdd(&h[b], 1u);
                }
            }
        }
    }
}

// one warp per term: terms with at least kSeedMinDf postings in the summed histogram get a slot and their 256 bins
__global__ void seed_import_kernel(const uint32_t *__restrict__ hist, uint32_t n_terms, uint32_t *__restrict__ slot_of_term,
                                   uint32_t *__restrict__ counter, uint32_t *__restrict__ cum)
{
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (t >= n_terms) return;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = hist[(size_t)t * 256 + lane * 8 + j];
        sum += v[j];
    }
    unsigned long long total = sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    uint32_t slot = kNoSeedSlot;
    if (total >= kSeedMinDf) {
        if (lane == 0) slot = atomicAdd(counter, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) cum[(size_t)slot * 256 + lane * 8 + j] = v[j];
    }
    if (lane == 0) slot_of_term[t] = slot;
}

__global__ void tile_keys_kernel(const uint64_t *__restrict__ term_offsets, uint32_t n_terms,
                                 const uint32_t *__restrict__ docids, const uint8_t *__restrict__ impacts,
                                 uint64_t n_post, const unsigned long long *__restrict__ first_zero,
                                 uint32_t doc_lo, uint32_t doc_hi, int tile_shift,
                                 uint64_t *__restrict__ keys, TileStats *__restrict__ stats)
{
    unsigned long long vis = 0;
    unsigned int mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = upper_bound_u64(term_offsets, (uint64_t)n_terms + 1, i) - 1;
        const uint32_t d = docids[i];
        uint64_t key = ~0ull;
        if (i < first_zero[t] && d >= doc_lo && d < doc_hi) {
            const uint32_t rel = d - doc_lo;
            const uint64_t tile = rel >> tile_shift;
            if (tile >= 0xFFFFu) {
                stats->bad_docid = 1;
            } else {
                const uint32_t local = rel & ((1u << tile_shift) - 1u);
                key = (tile << kTkTileShift) | ((uint64_t)t << kTkTermShift) | ((uint64_t)local_to_field(local) << kTkLocalShift) |
                      impacts[i];
                ++vis;
                if (d + 1 > mx) mx = d + 1;  // d < doc_hi <= 2^32-1
            }
        }
        keys[i] = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vis += __shfl_xor_sync(0xffffffffu, vis, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane_id() == 0) {
        if (vis) atomicAdd(&stats->n_visible, vis);
        if (mx) atomicMax(&stats->max_docid_plus1, mx);
    }
}

// Doc-major input (a collection as the indexer wrote it, create.py:33-35 order): one warp per document. The input is
// already ordered by tile and, inside a tile, by document — a stable sort of every tile on (term, parity) is all
// that is left to do. Postings with impact 0 are hidden (create.py writes them, inverted_index.py:50-51 never reads
// past the first one of an impact-sorted list, so none is ever visible); so are term ids outside the vocabulary.
__global__ void tile_keys_docmajor_kernel(const uint32_t *__restrict__ term_ids, const uint8_t *__restrict__ impacts,
                                          const uint64_t *__restrict__ doc_offsets, uint64_t n_docs, uint32_t n_terms,
                                          int tile_shift, uint64_t *__restrict__ keys, TileStats *__restrict__ stats)
{
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long vis = 0;
    for (uint64_t doc = warp; doc < n_docs; doc += n_warps) {
        const uint64_t lo = doc_offsets[doc], hi = doc_offsets[doc + 1];
        const uint64_t tile = doc >> tile_shift;
        const uint32_t field = local_to_field((uint32_t)doc & ((1u << tile_shift) - 1u));
        for (uint64_t i = lo + lane_id(); i < hi; i += 32) {
            uint32_t t = term_ids[i];
            const uint32_t v = impacts[i];
            if (t >= n_terms) { t = n_terms; stats->bad_docid = 2; }   // reported as an out-of-vocabulary term id
            else if (v == 0) t = n_terms;
            else ++vis;
            keys[i] = (tile << kTkTileShift) | ((uint64_t)t << kTkTermShift) | ((uint64_t)field << kTkLocalShift) | v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vis += __shfl_xor_sync(0xffffffffu, vis, o);
    if (lane_id() == 0 && vis) atomicAdd(&stats->n_visible, vis);
}

// key index at which every tile starts (n_tiles + 1 entries)
__global__ void tile_first_keys_kernel(const uint64_t *__restrict__ doc_offsets, uint64_t n_docs, int tile_shift,
                                       uint32_t n_tiles, uint64_t *__restrict__ first_key)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t <= n_tiles; t += gridDim.x * blockDim.x)
        first_key[t] = doc_offsets[min((uint64_t)t << tile_shift, n_docs)];
}

__global__ void seed_slots_from_df_kernel(const unsigned long long *__restrict__ df, uint32_t n_terms,
                                          uint32_t *__restrict__ slot_of_term, uint32_t *__restrict__ counter)
{
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_terms; t += gridDim.x * blockDim.x)
        slot_of_term[t] = df[t] >= kSeedMinDf ? atomicAdd(counter, 1u) : kNoSeedSlot;
}

__device__ __forceinline__ uint64_t seg_index(uint64_t key, uint32_t n_terms)
{
    const uint64_t tile = key >> kTkTileShift;
    const uint64_t term = (key >> kTkTermShift) & 0xFFFFFFu;
    return tile * n_terms + term;
}

// `seg_dup[s]` is set when a segment holds the same document twice (possible in hand-made CSR or
// a model that lists a term twice): such a segment cannot be stored densely (one byte per doc).
// `seg_odd[s]` (pre-set to 0xFFFFFFFF) receives the index of the segment's first odd document.
__global__ void seg_bounds_kernel(const uint64_t *__restrict__ ka, const uint64_t *__restrict__ kb,
                                  const uint32_t *__restrict__ cur, uint64_t n_keys, uint32_t n_terms,
                                  uint32_t *__restrict__ seg_begin, uint32_t *__restrict__ seg_end,
                                  uint32_t *__restrict__ seg_odd, uint32_t *__restrict__ seg_dup)
{
    constexpr int kParityBit = kTkLocalShift + 15;
    const uint64_t *__restrict__ keys = rs_result(ka, kb, cur);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += (uint64_t)gridDim.x * blockDim.x) {
        if (((keys[i] >> kTkTermShift) & 0xFFFFFFu) >= n_terms) continue;  // hidden
        const uint64_t id = keys[i] >> kTkTermShift;
        const bool starts = i == 0 || (keys[i - 1] >> kTkTermShift) != id;
        if (i && (keys[i - 1] >> kTkLocalShift) == (keys[i] >> kTkLocalShift)) seg_dup[seg_index(keys[i], n_terms)] = 1u;
        if (((keys[i] >> kParityBit) & 1u) && (starts || !((keys[i - 1] >> kParityBit) & 1u)))
            seg_odd[seg_index(keys[i], n_terms)] = (uint32_t)i;
        if (starts) seg_begin[seg_index(keys[i], n_terms)] = (uint32_t)i;
        if (i == n_keys - 1 || (keys[i + 1] >> kTkTermShift) != id) seg_end[seg_index(keys[i], n_terms)] = (uint32_t)(i + 1);
    }
}

// per (tile, term): choose dense (one byte per doc of the tile) or sparse (u32 per posting) storage.
// On entry size16[s] holds the duplicate flag written by seg_bounds_kernel.
__global__ void seg_size_kernel(const uint32_t *__restrict__ seg_begin, const uint32_t *__restrict__ seg_end,
                                const uint32_t *__restrict__ seg_odd, uint64_t n_segs, uint32_t n_terms, uint32_t tile_docs, uint32_t dense_ratio,
                                uint32_t *__restrict__ size16, uint32_t *__restrict__ n_flag,
                                unsigned long long *__restrict__ df, TileStats *__restrict__ stats)
{
    unsigned long long nd = 0, ns = 0, ndp = 0;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t n = seg_end[s] - seg_begin[s];
        uint32_t sz = 0, nf = 0;
        if (n) {
            if (size16[s]) stats->n_dup_segments = 1;  // benign race: everybody writes 1
            const bool dense = dense_ratio != 0xFFFFFFFFu && (uint64_t)n * dense_ratio >= tile_docs && size16[s] == 0;
            if (dense) {
                sz = tile_docs >> kDenseUnitShift;  // one byte per document of the tile
                nf = n | kDenseFlag;
            } else {  // even documents first, then odd ones, each padded to whole 16-byte units
                const uint32_t n_even = seg_odd[s] == 0xFFFFFFFFu ? n : seg_odd[s] - seg_begin[s];
                const uint32_t ue = (n_even + 3u) / 4u, uo = (n - n_even + 3u) / 4u;
                sz = ue + uo;
                nf = (ue << 16) | sz;
            }
            if (dense) { ++nd; ndp += n; } else ++ns;
            atomicAdd(&df[s % n_terms], (unsigned long long)n);
        }
        size16[s] = sz;
        n_flag[s] = nf;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nd += __shfl_xor_sync(0xffffffffu, nd, o);
        ns += __shfl_xor_sync(0xffffffffu, ns, o);
        ndp += __shfl_xor_sync(0xffffffffu, ndp, o);
    }
    if (lane_id() == 0) {
        if (nd) atomicAdd(&stats->n_dense_segments, nd);
        if (ns) atomicAdd(&stats->n_sparse_segments, ns);
        if (ndp) atomicAdd(&stats->n_dense_postings, ndp);
    }
}

__global__ void narrow_u8_kernel(const uint32_t *__restrict__ in, uint64_t n, uint8_t *__restrict__ out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)min(in[i], 255u);
}

__global__ void seg_desc_kernel(const uint32_t *__restrict__ off16, const uint32_t *__restrict__ n_flag, uint64_t n_segs,
                                SegDesc *__restrict__ desc)
{
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += (uint64_t)gridDim.x * blockDim.x)
        desc[s] = SegDesc{off16[s], n_flag[s]};
}

// dense segment: one impact byte per document of the tile, in the unit order dense_pass16 reads.
// sparse posting word = impact << 16 | byte offset of the accumulator word ((local >> 1) * 4); the
// parity of the document is implied by the unit the word sits in.
// seed_slot / seed_hist (both nullable): while every posting passes through, the impact histograms of the frequent
// terms are counted too (doc-major build; the term-major build counts them from its impact-sorted lists instead)
__global__ void fill_payload_kernel(const uint64_t *__restrict__ ka, const uint64_t *__restrict__ kb,
                                    const uint32_t *__restrict__ cur, uint64_t n_keys, uint32_t n_terms,
                                    const SegDesc *__restrict__ desc, const uint32_t *__restrict__ seg_begin,
                                    const uint32_t *__restrict__ seg_odd, uint32_t tile_docs, uint8_t *__restrict__ payload,
                                    const uint32_t *__restrict__ seed_slot, uint32_t *__restrict__ seed_hist,
                                    uint32_t *__restrict__ seg_max32)
{
    const uint64_t *__restrict__ keys = rs_result(ka, kb, cur);
    const uint64_t n_round = (n_keys + 31) & ~31ull;  // whole warps stay together for the warp-wide maximum below
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = i < n_keys ? keys[i] : ~0ull;
        const uint32_t term = (uint32_t)(key >> kTkTermShift) & 0xFFFFFFu;
        if (seg_max32) {  // largest impact per (tile, term): one atomic per warp while the warp stays inside one segment
            const uint64_t id = key >> kTkTermShift;
            const bool uniform = __all_sync(0xffffffffu, id == __shfl_sync(0xffffffffu, id, 0));
            const uint32_t imp = term < n_terms ? ((uint32_t)key & 0xFFu) : 0u;
            if (uniform) {
                const uint32_t m = __reduce_max_sync(0xffffffffu, imp);
                if (lane_id() == 0 && term < n_terms) atomicMax(&seg_max32[seg_index(key, n_terms)], m);
            } else if (term < n_terms) {
                atomicMax(&seg_max32[seg_index(key, n_terms)], imp);
            }
        }
        if (term >= n_terms) continue;  // hidden
        if (seed_slot) {
            const uint32_t slot = seed_slot[term];
            if (slot != kNoSeedSlot) atomicAdd(&seed_hist[(size_t)slot * 256 + ((uint32_t)key & 0xFFu)], 1u);
        }
        const uint64_t s = seg_index(key, n_terms);
        const SegDesc d = desc[s];
        const uint32_t field = (uint32_t)(key >> kTkLocalShift) & 0xFFFFu;
        const uint32_t imp = (uint32_t)key & 0xFFu;
        if (d.n_flag & kDenseFlag) {
            // unit u = documents 8u..8u+7 then 8(u+H)..8(u+H)+7, H = tile_docs / 16 (see dense_pass16)
            const uint32_t local = field_to_local(field), g = local >> 3, H = tile_docs >> 4;
            payload[(size_t)d.off16 * 16 + (size_t)(g < H ? g : g - H) * 16 + (g < H ? 0u : 8u) + (local & 7u)] = (uint8_t)imp;
        } else {
            const uint32_t word = (imp << 16) | ((field & 0x7FFFu) << 2);
            const size_t slot = (field >> 15) ? (size_t)(d.n_flag >> 16) * 4 + (i - seg_odd[s]) : (i - seg_begin[s]);
            reinterpret_cast<uint32_t *>(payload)[(size_t)d.off16 * 4 + slot] = word;
        }
    }
}

// Sparse segments are added with shared-memory atomics, four per thread and 128-bit load: instruction i of a warp touches
// the words 4L + i of the lanes' units. In docid order their banks are random (birthday collisions: ~2 wavefronts per
// instruction, 20 % of the score kernel's shared-memory traffic). The order of postings inside a parity part is free (a
// sum), so every run of 32 units (128 words) is re-ordered BY BANK: with four words per bank, word 4L + i lands in
// bank L and the instruction is conflict-free. One warp per (tile, term) segment; counting sort over 32 bins.
constexpr int kBankSortWarps = 8;

__global__ void __launch_bounds__(kBankSortWarps * 32) sparse_bank_sort_kernel(const SegDesc *__restrict__ desc, uint64_t n_segs,
                                                                            uint8_t *__restrict__ payload)
{
    __shared__ uint32_t s_cnt[kBankSortWarps][32];
    __shared__ uint32_t s_words[kBankSortWarps][128];
    const uint32_t lane = lane_id(), w = threadIdx.x >> 5;
    uint4 *payload4 = reinterpret_cast<uint4 *>(payload);
    for (uint64_t seg = (uint64_t)blockIdx.x * kBankSortWarps + w; seg < n_segs; seg += (uint64_t)gridDim.x * kBankSortWarps) {
        const SegDesc d = desc[seg];
        if ((d.n_flag & kDenseFlag) || d.n_flag == 0) continue;
        const uint32_t total = d.n_flag & 0xFFFFu, even = d.n_flag >> 16;
        for (int part = 0; part < 2; ++part) {
            const uint32_t lo = part ? even : 0u, hi = part ? total : even;
            for (uint32_t u0 = lo; u0 < hi; u0 += 32) {
                const uint32_t n_units = min(32u, hi - u0);
                if (n_units < 2) continue;  // a single unit is one thread's four atomics, issued one after the other
                s_cnt[w][lane] = 0;
                __syncwarp();
                uint4 v = make_uint4(0, 0, 0, 0);
                uint32_t r[4] = {0, 0, 0, 0};
                if (lane < n_units) {
                    v = payload4[(size_t)d.off16 + u0 + lane];
                    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) r[i] = atomicAdd(&s_cnt[w][(x[i] >> 2) & 31u], 1u);  // any order inside a bank
                }
                __syncwarp();
                const uint32_t c = s_cnt[w][lane];
                uint32_t incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (unsigned)o) incl += up;
                }
                __syncwarp();
                s_cnt[w][lane] = incl - c;  // first slot of the bank
                __syncwarp();
                if (lane < n_units) {
                    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) s_words[w][s_cnt[w][(x[i] >> 2) & 31u] + r[i]] = x[i];
                }
                __syncwarp();
                if (lane < n_units)
                    payload4[(size_t)d.off16 + u0 + lane] = make_uint4(s_words[w][4 * lane], s_words[w][4 * lane + 1],
                                                                       s_words[w][4 * lane + 2], s_words[w][4 * lane + 3]);
                __syncwarp();
            }
        }
    }
}

// raw .dat image -> docids / impacts arrays (records may start at any byte offset)
__global__ void decode_dat_kernel(const uint8_t *__restrict__ dat, const uint64_t *__restrict__ term_offsets,
                                  const uint64_t *__restrict__ term_start_byte, uint32_t n_terms, uint64_t n_post,
                                  uint32_t *__restrict__ docids, uint8_t *__restrict__ impacts)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_post; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = upper_bound_u64(term_offsets, (uint64_t)n_terms + 1, i) - 1;
        const uint8_t *p = dat + term_start_byte[t] + 5 * (i - term_offsets[t]);
        docids[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        impacts[i] = p[4];
    }
}

}  // namespace di
