// search.cuh — K3 (term-at-a-time scoring of a query batch against one document tile),
// K4 (exact deterministic top-k) and K5 (cross-shard merge).
//
// Replaces InvertedIndex.score (inverted_index.py:55-62) and the scoring loop of
// SparseSearch.search (nano_beir_evaluator.py:113-133). One CTA owns one (query, tile) work
// item: the tile's accumulators live in shared memory, postings stream in with 128-bit loads,
// and only documents that can still reach the query's top-k leave the SM.
//
// Launch structure: one score_tile_kernel launch per tile, grid = queries. Launching tile by
// tile keeps every CTA of a launch on the SAME tile, so a tile's postings are read from HBM
// once and then served from L2 to all queries of the batch; and it makes each query's
// candidate list single-writer (exactly one CTA per query per launch), so the running
// threshold needs no inter-CTA protocol.
#pragma once

#include "build.cuh"
#include "common.cuh"
#include "scan_sort.cuh"

namespace di {

constexpr int kScoreThreads = 256;
constexpr int kMaxSeg = 32;            // query terms handled per round inside a work item
constexpr int kSparseUnroll = 4;       // independent 128-bit posting loads in flight per thread
constexpr uint32_t kSortSmemKeys = 4096;

struct SearchArgs {
    const SegDesc *desc;        // [n_tiles][n_terms]
    const uint8_t *payload;
    const uint32_t *q_terms;
    const uint64_t *q_offsets;  // already offset to the first query of the batch
    uint64_t *cand;             // [n_queries][cap] candidate keys (unsorted)
    uint32_t *cnt;              // [n_queries] live candidates
    uint64_t *theta;            // [n_queries] lower bound on the k-th best key (0 = none yet)
    uint32_t n_terms, tile_docs, tile_shift, doc_lo;
    uint32_t cap, c0, k;
    int top_shift;              // highest radix-select digit that can be non-zero
};

// ---------------------------------------------------------------------------- exact selection
// k-th largest of n unique 64-bit keys (n >= k >= 1): MSB-first radix select, 8 bits per pass.
__device__ uint64_t block_select_kth(const uint64_t *keys, uint32_t n, uint32_t k, int top_shift,
                                     uint32_t *s_hist /*256*/, uint32_t *s_tmp /*2*/)
{
    uint64_t prefix = 0, mask = 0;
    uint32_t remaining = k;
    for (int shift = top_shift; shift >= 0; shift -= 8) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = keys[i];
            if ((key & mask) == prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const unsigned lane = threadIdx.x;
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // lane L owns bins 255-8L .. 248-8L, descending
                c[j] = s_hist[255 - 8 * lane - j];
                sum += c[j];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (unsigned)o) incl += t;
            }
            const uint32_t excl = incl - sum;
            if (excl < remaining && remaining <= incl) {
                uint32_t r = remaining - excl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (r <= c[j]) {
                        s_tmp[0] = 255 - 8 * lane - j;
                        s_tmp[1] = r;
                        break;
                    }
                    r -= c[j];
                }
            }
        }
        __syncthreads();
        prefix |= (uint64_t)s_tmp[0] << shift;
        mask |= 0xFFull << shift;
        remaining = s_tmp[1];
        __syncthreads();
    }
    return prefix;
}

// keeps keys >= theta, in place, order preserved; returns how many were kept
__device__ uint32_t block_compact_ge(uint64_t *keys, uint32_t n, uint64_t theta, uint32_t *s_scan /*33*/)
{
    uint32_t out = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t key = i < n ? keys[i] : 0ull;
        const uint32_t flag = (i < n && key >= theta) ? 1u : 0u;
        uint32_t total;
        const uint32_t pos = block_exclusive_scan(flag, s_scan, total);  // barriers inside: loads are done
        if (flag) keys[out + pos] = key;                                 // out + pos <= i
        out += total;
        __syncthreads();
    }
    return out;
}

template <typename Ptr>
__device__ void bitonic_sort_desc(Ptr a, uint32_t n_pow2)
{
    for (uint32_t k2 = 2; k2 <= n_pow2; k2 <<= 1) {
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t x = a[i], y = a[ixj];
                    const bool desc_block = (i & k2) == 0;
                    if (desc_block ? (x < y) : (x > y)) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------- K3: score one tile
template <bool ACC32>
__global__ void __launch_bounds__(kScoreThreads, ACC32 ? 1 : 3) score_tile_kernel(SearchArgs p, uint32_t tile)
{
    extern __shared__ uint4 s_acc4[];  // tile accumulators: u16 pairs (ACC16) or u32 (ACC32)
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(s_acc4);
    __shared__ uint32_t s_doff[kMaxSeg];       // dense segments of this round: payload offset (16 B units)
    __shared__ uint32_t s_soff[kMaxSeg];       // sparse segments: payload offset
    __shared__ uint32_t s_sunits[kMaxSeg];     // sparse segments: length in 16 B units (4 postings)
    __shared__ uint32_t s_spref[kMaxSeg + 1];  // exclusive prefix of s_sunits
    __shared__ uint32_t s_nd, s_ns, s_emit;
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tmp[2];

    const uint32_t tid = threadIdx.x;
    const uint32_t q = blockIdx.x;
    const uint64_t qb = p.q_offsets[q], qe = p.q_offsets[q + 1];
    const uint32_t T = p.tile_docs;
    const SegDesc *__restrict__ desc = p.desc + (uint64_t)tile * p.n_terms;
    const uint4 *__restrict__ payload4 = reinterpret_cast<const uint4 *>(p.payload);

    bool first = true, touched = false;
    for (uint64_t r0 = qb; first || r0 < qe; r0 += kMaxSeg) {
        // ---- look up this round's (term, tile) segments
        if (tid == 0) { s_nd = 0; s_ns = 0; }
        __syncthreads();
        if (tid < kMaxSeg && r0 + tid < qe) {
            const uint32_t t = p.q_terms[r0 + tid];
            if (t < p.n_terms) {  // DI_OOV_TERM and anything out of range: no postings
                const SegDesc d = desc[t];
                const uint32_t n = d.n_flag & ~kDenseFlag;
                if (n) {
                    if (d.n_flag & kDenseFlag) {
                        s_doff[atomicAdd(&s_nd, 1u)] = d.off16;
                    } else {
                        const uint32_t j = atomicAdd(&s_ns, 1u);
                        s_soff[j] = d.off16;
                        s_sunits[j] = (n + 3u) >> 2;
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t nd = s_nd, ns = s_ns;
        if (first && r0 + kMaxSeg >= qe && nd + ns == 0) return;  // query has no posting in this tile
        touched = touched || (nd + ns) != 0;
        if (tid == 0) {
            uint32_t run = 0;
            for (uint32_t j = 0; j < ns; ++j) { s_spref[j] = run; run += s_sunits[j]; }
            s_spref[ns] = run;
        }

        // ---- dense segments: u8 impact per document of the tile, summed in registers; the first
        //      round STORES the sums (this is also what zeroes the accumulators)
        if (first || nd) {
            const uint32_t groups = T >> 4;  // 16 documents per 128-bit load
            for (uint32_t g = tid; g < groups; g += kScoreThreads) {
                if (!ACC32) {
                    uint32_t a[8];
                    if (first) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) a[i] = 0;
                    } else {
                        const uint4 lo = s_acc4[2 * g], hi = s_acc4[2 * g + 1];
                        a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w;
                        a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
                    }
                    for (uint32_t j0 = 0; j0 < nd; j0 += 4) {
                        uint4 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            v[u] = (j0 + u < nd) ? ldg_stream_v4(payload4 + s_doff[j0 + u] + g) : make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {  // two u16 lanes per word, no carry between them
                            a[0] += __byte_perm(v[u].x, 0, 0x4140); a[1] += __byte_perm(v[u].x, 0, 0x4342);
                            a[2] += __byte_perm(v[u].y, 0, 0x4140); a[3] += __byte_perm(v[u].y, 0, 0x4342);
                            a[4] += __byte_perm(v[u].z, 0, 0x4140); a[5] += __byte_perm(v[u].z, 0, 0x4342);
                            a[6] += __byte_perm(v[u].w, 0, 0x4140); a[7] += __byte_perm(v[u].w, 0, 0x4342);
                        }
                    }
                    s_acc4[2 * g] = make_uint4(a[0], a[1], a[2], a[3]);
                    s_acc4[2 * g + 1] = make_uint4(a[4], a[5], a[6], a[7]);
                } else {
                    uint32_t a[16];
                    if (first) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) a[i] = 0;
                    } else {
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const uint4 x = s_acc4[4 * g + w];
                            a[4 * w] = x.x; a[4 * w + 1] = x.y; a[4 * w + 2] = x.z; a[4 * w + 3] = x.w;
                        }
                    }
                    for (uint32_t j = 0; j < nd; ++j) {
                        const uint4 v = ldg_stream_v4(payload4 + s_doff[j] + g);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            a[4 * i] += w[i] & 255u; a[4 * i + 1] += (w[i] >> 8) & 255u;
                            a[4 * i + 2] += (w[i] >> 16) & 255u; a[4 * i + 3] += w[i] >> 24;
                        }
                    }
#pragma unroll
                    for (int w = 0; w < 4; ++w)
                        s_acc4[4 * g + w] = make_uint4(a[4 * w], a[4 * w + 1], a[4 * w + 2], a[4 * w + 3]);
                }
            }
        }
        __syncthreads();

        // ---- sparse segments: u32 postings (impact << 16 | local docid), all segments of the round
        //      flattened into one index space so that short lists do not idle the CTA
        const uint32_t total = s_spref[ns];
        uint32_t seg = 0;
        for (uint32_t u0 = tid; u0 < total; u0 += kSparseUnroll * kScoreThreads) {
            uint4 v[kSparseUnroll];
#pragma unroll
            for (int j = 0; j < kSparseUnroll; ++j) {
                const uint32_t u = u0 + j * kScoreThreads;
                if (u < total) {
                    while (u >= s_spref[seg + 1]) ++seg;
                    v[j] = ldg_stream_v4(payload4 + s_soff[seg] + (u - s_spref[seg]));
                } else {
                    v[j] = make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int j = 0; j < kSparseUnroll; ++j) {
                const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t imp = w[i] >> 16;
                    if (imp) {  // padding words are 0
                        if (!ACC32)
                            atomicAdd(&s_acc[(w[i] & 0xFFFFu) >> 1], imp << ((w[i] & 1u) << 4));
                        else
                            atomicAdd(&s_acc[w[i] & 0xFFFFu], imp);
                    }
                }
            }
        }
        __syncthreads();
        first = false;
    }
    if (!touched) return;

    // ---- K4, tile-local part: documents whose key can still enter the top-k go to the query's
    //      candidate list. This CTA is the only writer of that list during this launch.
    const uint64_t theta = p.theta[q];
    const uint32_t cnt0 = p.cnt[q];
    uint32_t ths = (uint32_t)(theta >> 32);
    if (ths == 0) ths = 1;  // score 0 = document not touched: never a result (inverted_index.py:58-62)
    uint64_t *__restrict__ cand = p.cand + (uint64_t)q * p.cap;
    const uint32_t doc_base = p.doc_lo + (tile << p.tile_shift);
    if (tid == 0) s_emit = 0;
    __syncthreads();
    if (!ACC32) {
        const uint32_t groups = T >> 3;  // 8 u16 accumulators per 128-bit shared load
        for (uint32_t g = tid; g < groups; g += kScoreThreads) {
            const uint4 x = s_acc4[g];
            const uint32_t m2 = __vmaxu2(__vmaxu2(x.x, x.y), __vmaxu2(x.z, x.w));
            if (max(m2 & 0xFFFFu, m2 >> 16) >= ths) {
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t s = (w[i >> 1] >> ((i & 1) << 4)) & 0xFFFFu;
                    if (s >= ths) {
                        const uint64_t key = make_key(s, doc_base + 8 * g + i);
                        if (key >= theta) cand[cnt0 + atomicAdd(&s_emit, 1u)] = key;
                    }
                }
            }
        }
    } else {
        const uint32_t groups = T >> 2;
        for (uint32_t g = tid; g < groups; g += kScoreThreads) {
            const uint4 x = s_acc4[g];
            if (max(max(x.x, x.y), max(x.z, x.w)) >= ths) {
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (w[i] >= ths) {
                        const uint64_t key = make_key(w[i], doc_base + 4 * g + i);
                        if (key >= theta) cand[cnt0 + atomicAdd(&s_emit, 1u)] = key;
                    }
                }
            }
        }
    }
    __syncthreads();
    uint32_t n = cnt0 + s_emit;  // <= c0 + tile_docs <= cap
    if (n > p.c0) {
        // too many live candidates: keep exactly the k best and raise the threshold to the k-th
        const uint64_t kth = block_select_kth(cand, n, p.k, p.top_shift, s_hist, s_tmp);
        n = block_compact_ge(cand, n, kth, s_scan);
        if (tid == 0) p.theta[q] = kth;
    }
    if (tid == 0) p.cnt[q] = n;
}

// ---------------------------------------------------------------------------- K4: final select + sort
// One CTA per query: cut the candidate list to the k best and sort them descending by key
// (score desc, docid asc). Sorting happens in shared memory when the list fits.
__global__ void __launch_bounds__(kScoreThreads) finalize_topk_kernel(uint64_t *cand_all, const uint32_t *cnt, uint32_t cap,
                                                                    uint32_t k, int top_shift, uint64_t *out_keys,
                                                                    uint32_t *out_counts)
{
    extern __shared__ uint64_t s_keys[];  // kSortSmemKeys
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tmp[2];
    const uint32_t q = blockIdx.x;
    uint64_t *cand = cand_all + (uint64_t)q * cap;
    uint32_t n = cnt[q];
    if (n > k) {
        const uint64_t kth = block_select_kth(cand, n, k, top_shift, s_hist, s_tmp);
        n = block_compact_ge(cand, n, kth, s_scan);
    }
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    uint64_t *out = out_keys + (uint64_t)q * k;
    if (np2 <= kSortSmemKeys) {
        for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x) s_keys[i] = i < n ? cand[i] : 0ull;
        __syncthreads();
        bitonic_sort_desc(s_keys, np2);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = s_keys[i];
    } else {  // cap >= np2 is guaranteed by the host
        for (uint32_t i = n + threadIdx.x; i < np2; i += blockDim.x) cand[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(cand, np2);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = cand[i];
    }
    if (threadIdx.x == 0) out_counts[q] = n;
}

__global__ void unpack_keys_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ docids,
                                   int32_t *__restrict__ scores)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        docids[i] = key_docid(k);
        scores[i] = (int32_t)key_score(k);
    }
}

// ---------------------------------------------------------------------------- K5: cross-shard merge
// keys_in: [n_shards][n_queries][k] (each row sorted, valid up to counts_in[shard][q]).
// Gathers the rows of one query into a scratch list; finalize_topk_kernel then selects + sorts.
__global__ void merge_gather_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ counts_in,
                                    uint32_t n_shards, uint32_t n_queries, uint32_t k, uint64_t *__restrict__ cand,
                                    uint32_t *__restrict__ cnt, uint32_t cap)
{
    const uint32_t q = blockIdx.x;
    uint32_t run = 0;
    for (uint32_t s = 0; s < n_shards; ++s) {
        const uint32_t c = counts_in[(uint64_t)s * n_queries + q];
        const uint64_t *src = keys_in + ((uint64_t)s * n_queries + q) * k;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) cand[(uint64_t)q * cap + run + i] = src[i];
        run += c;
    }
    if (threadIdx.x == 0) cnt[q] = run;
}

}  // namespace di
