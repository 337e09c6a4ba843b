// score_tile.cuh — K3: term-at-a-time scoring of every query of a batch against ONE document
// tile, plus the tile-local half of K4 (candidate emission against the query's running threshold).
//
// Replaces the two hot loops of InvertedIndex.score (inverted_index.py:57-60: read the term's
// postings, scores[doc] += impact) and of SparseSearch.search (nano_beir_evaluator.py:118-121).
//
// One CTA = one work item = one query against kTilesPerItem adjacent tiles, one tile after the other. The shared-memory
// accumulators are ZERO when a tile starts (invariant of the 16-bit form: every pass that reads an accumulator word
// leaves a zero behind).
//   phase 1  sparse segments (u32 postings) of all the query's terms are flattened into one index
//            space, streamed with 128-bit loads and added with shared-memory atomics;
//   phase 2  ONE fused pass over the tile: dense segments (one impact byte per document, 128-bit loads)
//            are widened and summed in registers on top of the accumulator words, the sums are tested
//            against the query's threshold while still in registers (a bit per 8-document group in a
//            per-thread mask), and the word is written back as zero — or as the sum for the rare group
//            that holds a candidate. No separate store / scan passes, no barrier between them;
//   phase 3  the hit groups are expanded with every lane busy and the surviving keys (score, ~docid)
//            appended to the query's candidate list in global memory. Every query starts from a PROVEN
//            threshold (build.cuh, threshold seeds).
// Accumulators are u16 pairs packed in 32-bit words (ACC32 = false; queries of <= 257 terms cannot
// overflow 16 bits) or u32 (ACC32 = true: long queries, zeroed at item start, separate scan pass).
// Per tile the 16-bit form synchronises the CTA three times (after the sparse atomics, after the fused pass —
// a uniform vote that ends the tile when no group holds a candidate — and after the hit expansion); the segment
// lookup of all the item's tiles is shared by the four warps. The kernel compiles to exactly 80 registers (six
// CTAs per SM): anything that keeps one more value alive across a phase has measured slower.
#pragma once

#include "build.cuh"
#include "common.cuh"
#include "select.cuh"

namespace di {

#ifndef DI_SCORE_THREADS
#define DI_SCORE_THREADS 128
#endif
#ifndef DI_SCORE_MIN_BLOCKS
#define DI_SCORE_MIN_BLOCKS 6
#endif
constexpr int kScoreThreads = DI_SCORE_THREADS;
constexpr int kMaxSeg = 32;        // query terms handled per round inside a work item
#ifndef DI_SPARSE_UNROLL
#define DI_SPARSE_UNROLL 2
#endif
constexpr int kSparseUnroll = DI_SPARSE_UNROLL;  // independent 128-bit posting loads in flight per thread
#ifndef DI_HIST_BINS
#define DI_HIST_BINS 512
#endif
constexpr int kHistBins = DI_HIST_BINS;  // score histogram of the tile-local pre-selection = slots of the hit-group list (multiple of 256)
#ifndef DI_TILES_PER_ITEM
#define DI_TILES_PER_ITEM 8
#endif
constexpr int kTilesPerItem = DI_TILES_PER_ITEM;  // adjacent tiles one work item covers
#ifndef DI_FUSE_MAX
#define DI_FUSE_MAX 6   // measured: 4 -> 28.19, 6 -> 28.15, 8 -> 30.4 ms per step (one unit per step leaves too few loads in flight)
#endif
// Barrier-lean tile loop of the 16-bit form (default): the hit / emit counters grow over the whole work item (every thread
// keeps their values at tile start in registers) instead of being reset behind a barrier per tile, and the barriers that
// only separated a tile from the next one are dropped: sparse -> fused pass -> hit expansion are the three that remain.
#if defined(DI_NO_LEAN_BARRIERS) || defined(DI_DENSE_TMA)
constexpr bool kLean = false;
#else
constexpr bool kLean = true;
#endif
constexpr int kFuseMax = DI_FUSE_MAX;  // dense segments the fused dense + threshold pass sums in registers (4 .. 8)
static_assert(kFuseMax >= 4 && kFuseMax <= 8, "fused pass handles 4 to 8 dense segments");
static_assert(kHistBins % 256 == 0 && kHistBins >= kSelectSmemWords, "s_hist doubles as the radix select's scratch (a 256-bin build corrupts it: measured)");
static_assert(kTilesPerItem >= 1 && kTilesPerItem * 336 + DI_HIST_BINS * 4 <= 4800,
              "segment lists + hit list must leave room for six CTAs of 32 KB accumulators per SM (static smem <= 5 KB)");

constexpr int kRecInlineTerms = 12;
struct __align__(64) QueryRec {   // one cache-line-friendly record per query of the batch
    uint32_t q;                   // query index inside the batch
    uint32_t n;                   // number of term occurrences
    uint64_t begin;               // first term in q_terms (for queries longer than the inline part)
    uint32_t terms[kRecInlineTerms];
};

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
    const QueryRec *recs;       // [n_queries] work-item records in launch order
    uint32_t *done;             // [lanes * n_queries] tiles completed per (lane, query) (persistent launch), or nullptr
    // Small batches cannot fill the GPU with one tile chain per query, so the tile range is cut into
    // `lanes` contiguous sub-ranges that run independently (own candidate list and threshold per
    // (lane, query), all per-query arrays are [lanes][n_queries]) and are merged like shards afterwards.
    uint32_t n_queries, n_tiles, lanes, tiles_per_lane;
    // exact tile skipping (MaxScore-style bound, optional): seg_max[tile][term] = largest impact of the term inside the
    // tile; a (query, tile) whose bounds add up to less than the query's threshold cannot hold a result and is not
    // scored at all. nullptr = off (also for an index in which a posting list names a document twice).
    const uint8_t *seg_max;
    unsigned long long *n_skipped;  // tiles skipped that way (diagnostic counter)
#ifdef DI_PROFILE_PHASES
    unsigned long long *prof;   // [n_tiles][8] cycles per phase, summed over the tile's work items (diagnostic build)
#endif
};

// Diagnostic build only (-DDI_PROFILE_PHASES, tools/phase_profile.sh): thread 0 adds the cycles since the
// previous mark to prof[tile][phase]. Compiles to nothing in the product build.
#ifdef DI_PROFILE_PHASES
#define DI_PROF_DECL long long prof_t = clock64()
#define DI_PROF_MARK(phase)                                                              \
    do {                                                                                 \
        if (threadIdx.x == 0 && p.prof) {                                                \
            const long long now = clock64();                                             \
            atomicAdd(p.prof + (size_t)tile * 8 + (phase), (unsigned long long)(now - prof_t)); \
            prof_t = now;                                                                \
        }                                                                                \
    } while (0)
#else
#define DI_PROF_DECL
#define DI_PROF_MARK(phase)
#endif

// Launch order of a batch: queries bucketed by log2 of their total posting count (sum of the document
// frequencies of their terms), heaviest bucket first; the order inside a bucket is arbitrary (results do not
// depend on it). Two kernels: the buckets are found by the whole grid (one thread per query: the df lookups are
// dependent global loads), the counting sort over 64 buckets and the 64-byte records by one block.
__global__ void query_bucket_kernel(const uint32_t *__restrict__ q_terms, const uint64_t *__restrict__ q_offsets,
                                    const unsigned long long *__restrict__ df, uint32_t n_terms, uint32_t n_queries,
                                    uint8_t *__restrict__ bucket)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_queries) return;
    unsigned long long cost = 0;
    for (uint64_t j = q_offsets[q]; j < q_offsets[q + 1]; ++j) {
        const uint32_t t = q_terms[j];
        if (t < n_terms) cost += df[t];
    }
    // two buckets per power of two, descending cost = ascending bucket
    const int msb = cost ? 63 - __clzll((long long)cost) : 0;
    const int half = msb ? (int)((cost >> (msb - 1)) & 1ull) : 0;
    const int b = 2 * msb + half;
    bucket[q] = (uint8_t)(63 - (b > 63 ? 63 : b));
}

__global__ void __launch_bounds__(1024) query_order_kernel(const uint32_t *__restrict__ q_terms,
                                                         const uint64_t *__restrict__ q_offsets,
                                                         const uint8_t *__restrict__ bucket, uint32_t n_queries,
                                                         QueryRec *__restrict__ recs)
{
    __shared__ uint32_t s_cursor[64];
    if (threadIdx.x < 64) s_cursor[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < n_queries; q += blockDim.x) atomicAdd(&s_cursor[bucket[q]], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < 64; ++b) {
            const uint32_t c = s_cursor[b];
            s_cursor[b] = run;
            run += c;
        }
    }
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < n_queries; q += blockDim.x) {
        QueryRec r;
        r.q = q;
        r.begin = q_offsets[q];
        r.n = (uint32_t)(q_offsets[q + 1] - r.begin);
        for (int i = 0; i < kRecInlineTerms; ++i) r.terms[i] = (uint32_t)i < r.n ? q_terms[r.begin + i] : DI_OOV_TERM;
        recs[atomicAdd(&s_cursor[bucket[q]], 1u)] = r;
    }
}

// ---- accumulator groups ----------------------------------------------------------------------
// Accumulators are handled in groups of one 128-bit word: 8 documents (u16 pairs) or 4 (u32).
template <bool ACC32> __device__ __forceinline__ constexpr int group_docs() { return ACC32 ? 4 : 8; }

template <bool ACC32> __device__ __forceinline__ uint32_t group_score(const uint4 &x, int i)
{
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    if (ACC32) return w[i];
    return (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu);
}

// does the group hold a document with score >= ths?  (tm2 = (ths - 1) in both u16 halves)
template <bool ACC32>
__device__ __forceinline__ bool group_hit(const uint4 &x, uint32_t ths, uint32_t tm2)
{
    if (!ACC32) {  // per-lane max of the four words (3-input SIMD max), one more max against the threshold
        const uint32_t m = __vmaxu2(__vmaxu2(x.x, x.y), __vmaxu2(x.z, x.w));
        return __vmaxu2(m, tm2) != tm2;
    }
    return max(max(x.x, x.y), max(x.z, x.w)) >= ths;
}

// ---- dense segments, 16 documents per 128-bit load ---------------------------------------------
// A dense segment stores one BYTE per document of the tile. 16-byte unit u holds documents 8u .. 8u+7 and
// 8(u+H) .. 8(u+H)+7 (H = tile_docs / 16 units), i.e. the accumulator words s_acc4[u] and s_acc4[u + H]: a
// warp's accesses stay contiguous (no bank conflicts) and a byte pair widens to a packed u16 accumulator word
// with one PRMT. Thread t always owns the units u = t (mod block size), whatever NB is, so consecutive passes
// over the same words need no barrier.
__device__ __forceinline__ void add_bytes8(uint4 &a, uint32_t lo, uint32_t hi)
{
    a.x += __byte_perm(lo, 0, 0x4140);
    a.y += __byte_perm(lo, 0, 0x4342);
    a.z += __byte_perm(hi, 0, 0x4140);
    a.w += __byte_perm(hi, 0, 0x4342);
}

// Adds NB dense segments on top of the accumulator words.
//   FUSE = false: plain read-modify-write (dense terms beyond the four of the fused pass).
//   FUSE = true : the sums are tested against the threshold in registers; a word goes back as ZERO unless its
//                 group holds a candidate (then the sums are kept for expand_hits). Returns the thread's hit
//                 mask: bit 2j (+1) = the group of its j-th unit in the lower (upper) half of the tile.
// FULL: units is a multiple of the step (U * threads), so no load or store needs a bounds predicate.
// DI_DENSE_TMA (experiment): segment 0 of the fused pass was copied into shared memory by the TMA unit (`stage`); the
// pass then reads it with LDS instead of LDG.
#ifdef DI_DENSE_TMA
constexpr bool kDenseTma = true;
#else
constexpr bool kDenseTma = false;
#endif
// dynamic shared memory of a score CTA: the accumulators, plus one staged dense segment in the TMA experiment
__host__ __device__ inline size_t score_smem_bytes(uint32_t tile_docs, bool acc32)
{
    return (size_t)tile_docs * (acc32 ? 4 : 2) + ((kDenseTma && !acc32) ? tile_docs : 0);
}

// units per step so that about 8 independent 128-bit loads are in flight per thread whatever the number of dense
// terms (a work item is latency-bound: few CTAs per SM, L2-resident postings)
template <int NB> __device__ __forceinline__ constexpr int dense_units_per_step() { return NB <= 1 ? 8 : (NB == 2 ? 4 : (NB <= 4 ? 2 : 1)); }

// DI_DENSE_PRELOAD (variant): the loads of the fused pass's FIRST step are issued before the sparse phase and wait in
// registers (`pre`), so that their latency overlaps the sparse phase's own load -> atomic -> barrier chain.
#ifdef DI_DENSE_PRELOAD
constexpr bool kDensePreload = true;
#else
constexpr bool kDensePreload = false;
#endif
#ifndef DI_PRE_WORDS
#define DI_PRE_WORDS 4
#endif
constexpr int kPreWords = DI_PRE_WORDS;  // 128-bit words held per thread (8 = the whole first step; more than 4 spill at 80 registers)
template <int NB> __device__ __forceinline__ constexpr int dense_pre_units()
{
    return NB == 0 ? 0 : (dense_units_per_step<NB>() * NB <= kPreWords ? dense_units_per_step<NB>() : kPreWords / NB);
}

template <int NB>
__device__ __forceinline__ void dense_preload(uint4 (&pre)[kPreWords], const uint4 *payload4, const uint32_t *doff)
{
#ifdef DI_DENSE_PREFETCH_L1   // variant: no registers held, the first step's lines are only asked into L1
    constexpr int U = dense_units_per_step<NB>();
#pragma unroll
    for (int s = 0; s < U; ++s)
#pragma unroll
        for (int u = 0; u < NB; ++u)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(payload4 + doff[u] + threadIdx.x + s * kScoreThreads));
#else
    constexpr int UP = dense_pre_units<NB>();
#pragma unroll
    for (int s = 0; s < UP; ++s)
#pragma unroll
        for (int u = 0; u < NB; ++u) pre[s * NB + u] = ldg_dense_v4(payload4 + doff[u] + threadIdx.x + s * kScoreThreads);
#endif
}

__device__ __forceinline__ void dense_preload_dispatch(int nb, uint4 (&pre)[kPreWords], const uint4 *payload4, const uint32_t *doff)
{
    switch (nb) {  // nb is uniform across the CTA
#if DI_FUSE_MAX >= 8
        case 8: dense_preload<8>(pre, payload4, doff); break;
#endif
#if DI_FUSE_MAX >= 7
        case 7: dense_preload<7>(pre, payload4, doff); break;
#endif
#if DI_FUSE_MAX >= 6
        case 6: dense_preload<6>(pre, payload4, doff); break;
#endif
#if DI_FUSE_MAX >= 5
        case 5: dense_preload<5>(pre, payload4, doff); break;
#endif
        case 4: dense_preload<4>(pre, payload4, doff); break;
        case 3: dense_preload<3>(pre, payload4, doff); break;
        case 2: dense_preload<2>(pre, payload4, doff); break;
        case 1: dense_preload<1>(pre, payload4, doff); break;
        default: break;
    }
}

// one step of a dense pass: the U units g0, g0 + threads, ... whose NB dense words are in v
template <int NB, int U, bool FUSE, bool FULL>
__device__ __forceinline__ void dense_consume(uint4 *s_acc4, const uint4 (&v)[U][NB > 0 ? NB : 1], uint32_t g0, uint32_t ord,
                                              uint32_t units, uint32_t tm2, uint32_t &mask)
{
#pragma unroll
    for (int s = 0; s < U; ++s) {
        const uint32_t g = g0 + s * kScoreThreads;
        if (FULL || g < units) {
            uint4 a = s_acc4[g];
            uint4 b = s_acc4[g + units];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                add_bytes8(a, v[s][u].x, v[s][u].y);
                add_bytes8(b, v[s][u].z, v[s][u].w);
            }
            if (!FUSE) {
                s_acc4[g] = a;
                s_acc4[g + units] = b;
            } else {
                const bool ha = group_hit<false>(a, 0, tm2), hb = group_hit<false>(b, 0, tm2);
#ifdef DI_STORE_SELECT   // round-2 form: eight SEL + two stores per unit
                s_acc4[g] = ha ? a : make_uint4(0, 0, 0, 0);
                s_acc4[g + units] = hb ? b : make_uint4(0, 0, 0, 0);
#else                    // unconditional zero store, then the (rare) predicated store of a hit group's sums: -0.15 ms per step
                s_acc4[g] = make_uint4(0, 0, 0, 0);
                s_acc4[g + units] = make_uint4(0, 0, 0, 0);
                if (ha) s_acc4[g] = a;
                if (hb) s_acc4[g + units] = b;
#endif
                mask |= ((ha ? 1u : 0u) | (hb ? 2u : 0u)) << (2 * (ord + s));
            }
        }
    }
}

template <int NB, bool FUSE, bool FULL>
__device__ __forceinline__ uint32_t dense_steps16(uint4 *s_acc4, const uint4 *const (&ptr)[kFuseMax], uint32_t units, uint32_t tm2,
                                                  const uint4 *stage, const uint4 (*pre)[kPreWords])
{
    constexpr int U = dense_units_per_step<NB>();
    uint32_t mask = 0, ord = 0;
    uint32_t g0 = threadIdx.x;
    if (kDensePreload && FUSE && FULL && dense_pre_units<NB>() > 0 && pre) {  // first step: (some of) the words are already in registers
        constexpr int UP = dense_pre_units<NB>();
        uint4 v[U][NB > 0 ? NB : 1];
#pragma unroll
        for (int s = 0; s < U; ++s)
#pragma unroll
            for (int u = 0; u < NB; ++u) v[s][u] = s < UP ? (*pre)[s * NB + u] : ldg_dense_v4(ptr[u] + g0 + s * kScoreThreads);
        dense_consume<NB, U, FUSE, FULL>(s_acc4, v, g0, ord, units, tm2, mask);
        g0 += U * kScoreThreads;
        ord += U;
    }
    for (; g0 < units; g0 += U * kScoreThreads, ord += U) {
        uint4 v[U][NB > 0 ? NB : 1];
#pragma unroll
        for (int s = 0; s < U; ++s) {
            const uint32_t g = g0 + s * kScoreThreads;
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                if (kDenseTma && FUSE && u == 0 && stage)
                    v[s][u] = (FULL || g < units) ? stage[g] : make_uint4(0, 0, 0, 0);
                else
                    v[s][u] = (FULL || g < units) ? ldg_dense_v4(ptr[u] + g) : make_uint4(0, 0, 0, 0);
            }
        }
        dense_consume<NB, U, FUSE, FULL>(s_acc4, v, g0, ord, units, tm2, mask);
    }
    return mask;
}

// tiles of >= 16 K documents (the default) have only full steps in the fused pass
__device__ __forceinline__ bool dense_full_steps(uint32_t units) { return units % (8 * kScoreThreads) == 0; }

template <int NB, bool FUSE>
__device__ __forceinline__ uint32_t dense_pass16(uint4 *s_acc4, const uint4 *payload4, const uint32_t *doff, uint32_t units,
                                                 uint32_t tm2, const uint4 *stage, const uint4 (*pre)[kPreWords])
{
    const uint4 *ptr[kFuseMax];
#pragma unroll
    for (int u = 0; u < kFuseMax; ++u) ptr[u] = payload4 + doff[u < NB ? u : 0];
    if (FUSE && dense_full_steps(units)) return dense_steps16<NB, FUSE, true>(s_acc4, ptr, units, tm2, stage, pre);
    return dense_steps16<NB, FUSE, false>(s_acc4, ptr, units, tm2, stage, nullptr);
}

template <bool FUSE>
__device__ __forceinline__ uint32_t dense_dispatch16(int nb, uint4 *s_acc4, const uint4 *payload4, const uint32_t *doff,
                                                     uint32_t units, uint32_t tm2, const uint4 *stage = nullptr,
                                                     const uint4 (*pre)[kPreWords] = nullptr)
{
    switch (nb) {  // nb is uniform across the CTA
#if DI_FUSE_MAX >= 8
        case 8: return dense_pass16<8, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
#endif
#if DI_FUSE_MAX >= 7
        case 7: return dense_pass16<7, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
#endif
#if DI_FUSE_MAX >= 6
        case 6: return dense_pass16<6, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
#endif
#if DI_FUSE_MAX >= 5
        case 5: return dense_pass16<5, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
#endif
        case 4: return dense_pass16<4, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
        case 3: return dense_pass16<3, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
        case 2: return dense_pass16<2, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
        case 1: return dense_pass16<1, FUSE>(s_acc4, payload4, doff, units, tm2, stage, pre);
        default: return FUSE ? dense_pass16<0, true>(s_acc4, payload4, doff, units, tm2, nullptr, nullptr) : 0u;
    }
}

// ACC32 twin (long queries only): plain read-modify-write loop; the unit's two 8-document halves go to
// accumulator words 2u, 2u+1 and 2(u+H), 2(u+H)+1
__device__ __forceinline__ void dense_pass32(uint4 *s_acc4, const uint4 *payload4, const uint32_t *doff, int nb,
                                             uint32_t units)
{
    for (uint32_t g = threadIdx.x; g < units; g += kScoreThreads) {
        uint32_t a[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint4 lo = s_acc4[2 * (g + h * units)], hi = s_acc4[2 * (g + h * units) + 1];
            a[8 * h + 0] = lo.x; a[8 * h + 1] = lo.y; a[8 * h + 2] = lo.z; a[8 * h + 3] = lo.w;
            a[8 * h + 4] = hi.x; a[8 * h + 5] = hi.y; a[8 * h + 6] = hi.z; a[8 * h + 7] = hi.w;
        }
        for (int j = 0; j < nb; ++j) {
            const uint4 v = ldg_stream_v4(payload4 + doff[j] + g);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] += (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            s_acc4[2 * (g + h * units)] = make_uint4(a[8 * h + 0], a[8 * h + 1], a[8 * h + 2], a[8 * h + 3]);
            s_acc4[2 * (g + h * units) + 1] = make_uint4(a[8 * h + 4], a[8 * h + 5], a[8 * h + 6], a[8 * h + 7]);
        }
    }
}

// ---- hit groups -> candidates ----------------------------------------------------------------
// slow path (list already holds more than kHistBins hit groups): one shared-memory atomic per candidate
template <bool ACC32>
__device__ __forceinline__ void emit_group(const uint4 &x, uint32_t g, uint32_t ths, uint32_t doc_base, uint64_t theta,
                                           uint64_t *cand, uint32_t cnt0, uint32_t *s_emit)
{
#pragma unroll
    for (int i = 0; i < group_docs<ACC32>(); ++i) {
        const uint32_t sc = group_score<ACC32>(x, i);
        if (sc >= ths) {
            const uint64_t key = make_key(sc, doc_base + group_docs<ACC32>() * g + i);
            if (key >= theta) cand[cnt0 + atomicAdd(s_emit, 1u)] = key;
        }
    }
}

// The fused pass leaves one hit mask per thread; the warp compacts them into the hit-group list s_hits with
// one prefix sum and one shared-memory atomic (first kHistBins groups are remembered; *s_nhits counts all).
// *s_nhits is never reset inside a work item (no barrier needed for that): `nhits_base` is its value when the tile started.
__device__ __forceinline__ void record_hits16(uint32_t mask, uint32_t units, uint32_t *s_hits, uint32_t *s_nhits, uint32_t nhits_base)
{
    if (!__any_sync(0xffffffffu, mask != 0)) return;  // the usual case once the query has a threshold
    const uint32_t lane = lane_id(), c = __popc(mask);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += up;
    }
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(s_nhits, incl) - nhits_base;
    uint32_t slot = __shfl_sync(0xffffffffu, base, 31) + incl - c;
    while (mask) {
        const uint32_t bit = (uint32_t)(__ffs(mask) - 1);
        mask &= mask - 1u;
        if (slot < (uint32_t)kHistBins) s_hits[slot] = threadIdx.x + (bit >> 1) * kScoreThreads + ((bit & 1u) ? units : 0u);
        ++slot;
    }
}

// One pass over the tile's accumulators (32-bit form, and the rare flooded 16-bit tile): groups holding a
// document with score >= ths are REMEMBERED in s_hits (first kHistBins of them; *s_nhits counts all). Emitting on
// the spot would run the key / compare / append sequence with one or two lanes active; expand_hits() does it
// with every lane busy. The hot loop only sets a bit per hit group in a per-thread register mask (no branch,
// no atomic); after every 32 steps the warp compacts its masks with one prefix sum and one atomic.
// INPLACE: groups beyond the kHistBins slots are emitted on the spot instead of being dropped.
// The 16-bit form keeps its invariant here too: whatever is not handed to expand_hits is zeroed.
template <bool ACC32, bool INPLACE>
__device__ __forceinline__ void scan_groups(uint4 *s_acc4, uint32_t T, uint32_t ths, uint32_t *s_hits,
                                            uint32_t *s_nhits, uint32_t doc_base, uint64_t theta, uint64_t *cand,
                                            uint32_t cnt0, uint32_t *s_emit)
{
    const uint32_t groups = T / group_docs<ACC32>();  // a multiple of 32: every lane of a warp runs the same steps
    const uint32_t tm = ths - 1u, tm2 = tm | (tm << 16);  // u16 half > tm  <=>  half >= ths
    const uint32_t lane = lane_id();
    for (uint32_t c0 = threadIdx.x; c0 < groups; c0 += 32 * kScoreThreads) {  // chunk = 32 steps of the block
        uint32_t mask = 0;
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 4) {
            if (c0 + i0 * kScoreThreads < groups) {  // warp-uniform
                uint4 x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t g = c0 + (i0 + i) * kScoreThreads;
                    x[i] = g < groups ? s_acc4[g] : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t g = c0 + (i0 + i) * kScoreThreads;
                    if (group_hit<ACC32>(x[i], ths, tm2)) mask |= 1u << (i0 + i);
                    else if (!ACC32 && g < groups) s_acc4[g] = make_uint4(0, 0, 0, 0);
                }
            }
        }
        if (!__any_sync(0xffffffffu, mask != 0)) continue;
        const uint32_t c = __popc(mask);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += up;
        }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(s_nhits, incl);
        uint32_t slot = __shfl_sync(0xffffffffu, base, 31) + incl - c;
        while (mask) {
            const uint32_t g = c0 + (uint32_t)(__ffs(mask) - 1) * kScoreThreads;
            mask &= mask - 1u;
            if (slot < (uint32_t)kHistBins) {
                s_hits[slot] = g;
            } else {
                if (INPLACE) emit_group<ACC32>(s_acc4[g], g, ths, doc_base, theta, cand, cnt0, s_emit);
                if (!ACC32) s_acc4[g] = make_uint4(0, 0, 0, 0);
            }
            ++slot;
        }
    }
}

// Appends the candidates of the remembered groups: one group per lane, ONE shared-memory atomic per warp
// (warp prefix sum of the per-lane counts), keys written to cand[cnt0 + ...] in arbitrary order. The 16-bit
// form zeroes each group behind itself (accumulator invariant).
template <bool ACC32>
__device__ __forceinline__ void expand_hits(uint4 *s_acc4, const uint32_t *s_hits, uint32_t n_hits, uint32_t ths,
                                            uint32_t doc_base, uint64_t theta, uint64_t *cand, uint32_t cnt0,
                                            uint32_t *s_emit)
{
    const uint32_t lane = lane_id();
    for (uint32_t j0 = threadIdx.x - lane; j0 < n_hits; j0 += kScoreThreads) {  // warp-uniform trip count
        const uint32_t j = j0 + lane;
        uint32_t flags = 0, g = 0;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (j < n_hits) {
            g = s_hits[j];
            x = s_acc4[g];
            if (!ACC32) s_acc4[g] = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int i = 0; i < group_docs<ACC32>(); ++i) {
                const uint32_t sc = group_score<ACC32>(x, i);
                if (sc >= ths && make_key(sc, doc_base + group_docs<ACC32>() * g + i) >= theta) flags |= 1u << i;
            }
        }
        const uint32_t c = __popc(flags);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += up;
        }
        uint32_t base = 0;
        if (lane == 31 && incl) base = atomicAdd(s_emit, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t pos = cnt0 + base + incl - c;
#pragma unroll
        for (int i = 0; i < group_docs<ACC32>(); ++i)
            if (flags & (1u << i)) cand[pos++] = make_key(group_score<ACC32>(x, i), doc_base + group_docs<ACC32>() * g + i);
    }
}

__device__ __forceinline__ void zero_words16(uint4 *s, uint32_t n16)
{
    for (uint32_t i = threadIdx.x; i < n16; i += kScoreThreads) s[i] = make_uint4(0, 0, 0, 0);
}

// Segment lists of one (query, tile, round): filled by warp 0, one query term per lane, ballot-compacted.
// (Kept small on purpose: six CTAs of 32 KB accumulators + this fit one SM only while the static part stays <= 5 KB.)
struct SegLists {
    uint32_t off[kMaxSeg];        // payload offsets (16 B units): dense segments from the front, sparse ones from the back
    uint16_t seven[kMaxSeg];      // sparse segments: units holding even documents (they come first)
    uint32_t spref[kMaxSeg + 1];  // exclusive prefix of the sparse segments' unit counts
    uint32_t nd, ns;
    uint32_t ub;                  // sum over the round's terms of their largest impact in this tile (kNoBound = unknown)
    __device__ __forceinline__ uint32_t soff(uint32_t j) const { return off[kMaxSeg - 1 - j]; }
};

// max_impact: this lane's term's largest impact in the tile, or kNoBound (warp-uniform) when no bound is known.
constexpr uint32_t kNoBound = 0xFFFFFFFFu;

__device__ __forceinline__ void fill_seg_lists(SegLists &L, SegDesc d, uint32_t lane, uint32_t max_impact)
{
    uint32_t ub = max_impact;
    if (max_impact != kNoBound) {  // warp-uniform branch
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ub += __shfl_xor_sync(0xffffffffu, ub, o);
    }
    if (lane == 0) L.ub = ub;
    const bool is_dense = (d.n_flag & kDenseFlag) != 0, is_sparse = !is_dense && d.n_flag != 0;
    const uint32_t bd = __ballot_sync(0xffffffffu, is_dense), bs = __ballot_sync(0xffffffffu, is_sparse);
    const uint32_t su = is_sparse ? (d.n_flag & 0xFFFFu) : 0u;
    uint32_t incl = su;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += up;
    }
    if (is_dense) L.off[__popc(bd & lanemask_lt())] = d.off16;
    if (is_sparse) {
        const uint32_t j = __popc(bs & lanemask_lt());
        L.off[kMaxSeg - 1 - j] = d.off16;
        L.seven[j] = (uint16_t)(d.n_flag >> 16);
        L.spref[j] = incl - su;
    }
    if (lane == 31) L.spref[__popc(bs)] = incl;  // total units
    if (lane == 0) { L.nd = __popc(bd); L.ns = __popc(bs); }
}

// DI_SPARSE_PREFETCH (variant): the sparse units the NEXT tile of the item will read are asked into L1 (no registers
// held) right after this tile's fused pass — between here and the next tile's sparse loop there are only the hit
// expansion and its barriers, no dense streaming that would evict them again.
__device__ __forceinline__ void sparse_prefetch_l1(const SegLists &L, const uint4 *payload4)
{
    const uint32_t ns = L.ns;
    if (!ns) return;
    const uint32_t total = L.spref[ns];
    uint32_t seg = 0, seg_lo = 0, seg_hi = L.spref[1], seg_off = L.soff(0);
    for (uint32_t u = threadIdx.x; u < total; u += kScoreThreads) {
        if (u >= seg_hi) {
            do { ++seg; seg_hi = L.spref[seg + 1]; } while (u >= seg_hi);
            seg_lo = L.spref[seg];
            seg_off = L.soff(seg);
        }
        asm volatile("prefetch.global.L1 [%0];" ::"l"(payload4 + seg_off + (u - seg_lo)));
    }
}

// One work item = one query against n_sub (1 .. kTilesPerItem) ADJACENT tiles starting at tile0; `slot` indexes the
// batch's launch-ordered query records. Several tiles per item divide the per-item fixed costs (claim, record read,
// hand-off) and put the descriptor loads of all the tiles in flight together; the query's threshold and count travel
// from tile to tile in registers. Returns whether the query's state was read (i.e. whether the hand-off from the
// previous item was observed).
template <bool ACC32, bool BOUNDS>
__device__ __forceinline__ bool score_item(const SearchArgs &p, uint32_t tile0, uint32_t n_sub, uint32_t slot, uint32_t lane,
                                           uint32_t step, uint32_t &tma_phase, uint64_t *mbar)
{
    static_assert(kMaxSeg == 32, "the segment lookup is one warp wide");
    extern __shared__ uint4 s_acc4[];  // tile accumulators
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(s_acc4);
    __shared__ SegLists s_seg[kTilesPerItem];  // one per tile of the item
    __shared__ uint32_t s_emit, s_ready, s_nhits, s_cut;
    __shared__ __align__(16) uint32_t s_hist[kHistBins];  // score histogram / hit-group list / radix-select scratch
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tmp[2];

    const uint32_t tid = threadIdx.x;
    // One 64-byte record per work item (heaviest queries first, so the tail is made of light ones)
    // holds everything needed to start: the dependent-load chain is record -> descriptor -> postings.
    const QueryRec *__restrict__ rec = p.recs + slot;
    const uint32_t q = rec->q;
    const uint32_t sq = lane * p.n_queries + q;  // this (lane, query)'s slot in the per-query state arrays
    const uint64_t qb = rec->begin, qe = rec->begin + rec->n;
    const uint32_t T = p.tile_docs, units = T >> kDenseUnitShift;
    // The query's running state (threshold, candidate list) is handed from item to item through
    // global memory; step s may start once step s-1 of the same query has published (done[q] >= s).
    // Items are dispatched in tile-major order, so this is almost always already true: probe now,
    // and only wait (before the first fused pass) in the rare case it is not.
    if (tid == 0) s_ready = p.done == nullptr || step == 0 || ld_flag_u32(p.done + sq) >= step;
    constexpr bool lean = kLean && !ACC32;
    if (lean && tid == 0) { s_emit = 0; s_nhits = 0; }  // ordered by the first tile's list barrier; never reset inside the item
    uint32_t emit_base = 0, nhits_base = 0;              // the counters' values when the current tile started (uniform)
    const uint4 *__restrict__ payload4 = reinterpret_cast<const uint4 *>(p.payload);

    // ---- first-round segment lookup of ALL the item's tiles, one query term per lane: warp w takes the tiles w, w + warps, ...
    //      (round 2: warp 0 did all of them — eight serial ballot / prefix-scan chains), a lane's descriptor loads in flight together
#ifdef DI_LOOKUP_ONE_WARP
    constexpr int kLookupWarps = 1;
#else
    constexpr int kLookupWarps = kScoreThreads / 32;
#endif
    constexpr int kTilesPerWarp = (kTilesPerItem + kLookupWarps - 1) / kLookupWarps;
    if (tid < kLookupWarps * 32) {
        const uint32_t ln = tid & 31u, w = tid >> 5;
        SegDesc d[kTilesPerWarp];
        uint32_t mx[kTilesPerWarp];   // this term's largest impact in each tile; without the table: unknown (no skipping)
        const bool bounded = BOUNDS && qe - qb <= kMaxSeg;
#pragma unroll
        for (int jj = 0; jj < kTilesPerWarp; ++jj) { d[jj] = SegDesc{0u, 0u}; mx[jj] = 0u; }
        if (qb + ln < qe) {
            const uint32_t t = ln < kRecInlineTerms ? rec->terms[ln] : p.q_terms[qb + ln];
            if (t < p.n_terms) {  // DI_OOV_TERM and anything out of range: no postings
#pragma unroll
                for (int jj = 0; jj < kTilesPerWarp; ++jj) {
                    const uint32_t j = w + jj * kLookupWarps;
                    if (j < n_sub) {
                        d[jj] = p.desc[(uint64_t)(tile0 + j) * p.n_terms + t];
                        if (bounded) mx[jj] = p.seg_max[(uint64_t)(tile0 + j) * p.n_terms + t];
                    }
                }
            }
        }
        if (!bounded) {
#pragma unroll
            for (int jj = 0; jj < kTilesPerWarp; ++jj) mx[jj] = kNoBound;
        }
#ifdef DI_L2_PREFETCH
        // experiment (profiles/README.md): the item's later tiles are needed a few microseconds from now — ask the TMA
        // unit to pull their segments from HBM into L2 meanwhile (cp.async.bulk.prefetch.L2)
#pragma unroll
        for (int jj = 0; jj < kTilesPerWarp; ++jj) {
            const uint32_t j = w + jj * kLookupWarps;
            if (j >= 1 && j < n_sub && d[jj].n_flag) {
                const uint32_t units16 = (d[jj].n_flag & kDenseFlag) ? (p.tile_docs >> kDenseUnitShift) : (d[jj].n_flag & 0xFFFFu);
                prefetch_l2_bulk(p.payload + (uint64_t)d[jj].off16 * 16, units16 * 16);
            }
        }
#endif
#pragma unroll
        for (int jj = 0; jj < kTilesPerWarp; ++jj) {
            const uint32_t j = w + jj * kLookupWarps;  // warp-uniform
            if (j < n_sub) fill_seg_lists(s_seg[j], d[jj], ln, mx[jj]);
        }
    }

    DI_PROF_DECL;
    uint64_t theta = 0;
    uint32_t cnt0 = 0;
    bool have_state = false, dirty = false;  // dirty: threshold or count changed
    for (uint32_t sub = 0; sub < n_sub; ++sub) {
        const uint32_t tile = tile0 + sub;
        SegLists &L = s_seg[sub];
        const SegDesc *__restrict__ desc = p.desc + (uint64_t)tile * p.n_terms;
        uint32_t f0 = 0, nf = 0;  // L.doff[f0 .. f0 + nf): dense segments left for the fused pass
        uint4 pre[kPreWords];            // DI_DENSE_PRELOAD: first step of the fused pass, loaded ahead of the sparse phase
        bool preloaded = false;
        bool first = true, touched = false, skip = false;
        for (uint64_t r0 = qb; first || r0 < qe; r0 += kMaxSeg) {
            if (!first) {
                __syncthreads();  // the previous round's readers of the segment lists are done
                if (tid < kMaxSeg) {
                    SegDesc d{0u, 0u};
                    if (r0 + tid < qe) {
                        const uint32_t t = p.q_terms[r0 + tid];
                        if (t < p.n_terms) d = desc[t];
                    }
                    fill_seg_lists(L, d, tid, kNoBound);  // later rounds: no bound
                }
            }
            if (ACC32 && first) zero_words16(s_acc4, T / 4);  // overlaps the descriptor loads of warp 0
            // the lists of ALL the item's tiles were published by the first tile's barrier
            if (!lean || !first || sub == 0) __syncthreads();
            const uint32_t nd = L.nd, ns = L.ns;
            const bool last = r0 + kMaxSeg >= qe;
            if (first && last && nd + ns == 0) { skip = true; break; }  // query has no posting in this tile
            if (!have_state && s_ready) {  // state already published: fetch it now (L2), it is needed only by the fused pass
                theta = ld_cg_u64(p.theta + sq);
                cnt0 = ld_cg_u32(p.cnt + sq);
                have_state = true;
            }
            if (BOUNDS && first && have_state && L.ub != kNoBound) {
                // Exact skip: no document of this tile can reach the threshold (every term adds at most its largest
                // impact in the tile). Same tie rule as below: a tie with a threshold document of an earlier tile loses.
                uint32_t need = (uint32_t)(theta >> 32);
                if (need == 0) need = 1;
                if (theta != 0 && key_docid(theta) < p.doc_lo + (tile << p.tile_shift)) ++need;
                if (L.ub < need) {
                    if (tid == 0 && p.n_skipped) atomicAdd(p.n_skipped, 1ull);
                    skip = true;
                    break;
                }
            }
            touched = touched || (nd + ns) != 0;
            DI_PROF_MARK(0);  // segment lookup

            // ---- dense segments of the fused pass (the last four .. kFuseMax of the last round)
            nf = (!ACC32 && last) ? min(nd, (uint32_t)kFuseMax) : 0u;
            f0 = nd - nf;
            if (kDensePreload && !ACC32) {
                preloaded = nf != 0 && ns != 0 && dense_full_steps(units);
                if (preloaded) dense_preload_dispatch((int)nf, pre, payload4, L.off + f0);
            }

            // ---- phase 1: sparse segments. word = impact << 16 | byte offset of the accumulator word;
            //      units [0, even) of a segment hold even documents, the rest odd ones.
            if (ns) {
                const uint32_t total = L.spref[ns];
                // the current segment's bounds live in registers and are re-read only when a unit crosses into the
                // next segment (segments are hundreds of units long for the terms queries actually use)
                uint32_t seg = 0, seg_lo = 0, seg_hi = L.spref[1], seg_even = L.seven[0], seg_off = L.soff(0);
                for (uint32_t u0 = tid; u0 < total; u0 += kSparseUnroll * kScoreThreads) {
                    uint4 v[kSparseUnroll];
                    bool odd[kSparseUnroll];
#pragma unroll
                    for (int j = 0; j < kSparseUnroll; ++j) {
                        const uint32_t u = u0 + j * kScoreThreads;
                        odd[j] = false;
                        if (u < total) {
                            if (u >= seg_hi) {
                                do { ++seg; seg_hi = L.spref[seg + 1]; } while (u >= seg_hi);
                                seg_lo = L.spref[seg];
                                seg_even = L.seven[seg];
                                seg_off = L.soff(seg);
                            }
                            const uint32_t lu = u - seg_lo;
                            odd[j] = lu >= seg_even;
                            v[j] = ldg_sparse_v4(payload4 + seg_off + lu);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < kSparseUnroll; ++j) {
                        if (u0 + j * kScoreThreads < total) {
                            const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                            if (!ACC32) {
                                char *base = reinterpret_cast<char *>(s_acc);
                                if (!odd[j]) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<uint32_t *>(base + (w[i] & 0xFFFFu)), w[i] >> 16);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 4; ++i)
                                        atomicAdd(reinterpret_cast<uint32_t *>(base + (w[i] & 0xFFFFu)), w[i] & 0xFFFF0000u);
                                }
                            } else {
                                const uint32_t par = odd[j] ? 1u : 0u;
#pragma unroll
                                for (int i = 0; i < 4; ++i) atomicAdd(&s_acc[((w[i] & 0xFFFFu) >> 1) | par], w[i] >> 16);
                            }
                        }
                    }
                }
                __syncthreads();  // every atomic has landed before a thread reads its own accumulator words
            }
            DI_PROF_MARK(2);  // sparse

            // ---- dense segments that do not go through the fused pass: plain read-modify-write
            if (kDenseTma && !ACC32 && nf && tid == 0) {  // experiment: the TMA unit stages the first fused segment meanwhile
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar, T);
                tma_load_1d(s_acc4 + T / 8, payload4 + L.off[f0], T, mbar);
            }
            if (!ACC32) {
                for (uint32_t j0 = 0; j0 < f0; j0 += kFuseMax)
                    dense_dispatch16<false>(f0 - j0 < (uint32_t)kFuseMax ? (int)(f0 - j0) : kFuseMax, s_acc4, payload4, L.off + j0, units, 0u);
            } else if (nd) {
                dense_pass32(s_acc4, payload4, L.off, (int)nd, units);
            }
            DI_PROF_MARK(1);  // dense (read-modify-write)
            first = false;
        }
        if (skip || !touched) continue;  // nothing was added: the accumulators are still zero

        // ---- the query's state. Between done[q] == step and the publish after this item, this CTA is
        //      the only reader and writer of the query's list.
        if (!have_state) {
            if (tid == 0) {
                uint32_t spins = 0;
                while (ld_flag_u32(p.done + sq) < step) {
                    __nanosleep(128);
                    if (++spins > (1u << 23)) __trap();  // > 1 s: a protocol bug must fail loudly, never hang the GPU
                }
            }
            __syncthreads();
            theta = ld_cg_u64(p.theta + sq);
            cnt0 = ld_cg_u32(p.cnt + sq);
            have_state = true;
        }
        uint32_t ths = (uint32_t)(theta >> 32);
        if (ths == 0) ths = 1;  // score 0 = document not touched: never a result (inverted_index.py:58-62)
        uint64_t *__restrict__ cand = p.cand + (uint64_t)sq * p.cap;
        const uint32_t doc_base = p.doc_lo + (tile << p.tile_shift);
        // A document that only TIES the threshold score needs docid <= the threshold's docid. Tiles are
        // visited in docid order, so the threshold's document lies in an earlier tile and every such tie in
        // this tile loses: compare against score + 1 and keep the (many) ties out of the slow path.
        if (theta != 0 && key_docid(theta) < doc_base) ++ths;
        if (!ACC32 && ths > 0x10000u) ths = 0x10000u;  // a caller-given bound no 16-bit sum can reach: nothing hits
        uint64_t theta_pre = 0;
        if (!lean) {
            if (tid == 0) { s_emit = 0; s_nhits = 0; }
            __syncthreads();
        }
        DI_PROF_MARK(3);  // state wait

        // ---- phase 2: fused dense + threshold pass (16-bit), or the plain scan (32-bit).
        //      s_hist doubles as the hit-group list; it is free between the pre-selection and the radix select
        uint32_t mask = 0;
#ifdef DI_SPARSE_PREFETCH_EARLY
        if (!ACC32 && sub + 1 < n_sub) sparse_prefetch_l1(s_seg[sub + 1], payload4);
#endif
        if (!ACC32) {
            const uint32_t tm = ths - 1u;
            const uint4 *stage = nullptr;
            if (kDenseTma && nf) {
                mbar_wait(mbar, tma_phase & 1u);
                ++tma_phase;
                stage = s_acc4 + T / 8;
            }
            mask = dense_dispatch16<true>((int)nf, s_acc4, payload4, L.off + f0, units, tm | (tm << 16), stage,
#ifdef DI_DENSE_PREFETCH_L1
                                                         nullptr);
#else
                                                         (kDensePreload && preloaded) ? &pre : nullptr);
#endif
            record_hits16(mask, units, s_hist, &s_nhits, nhits_base);
#ifdef DI_SPARSE_PREFETCH
            if (sub + 1 < n_sub) sparse_prefetch_l1(s_seg[sub + 1], payload4);
#endif
        } else {
            scan_groups<ACC32, false>(s_acc4, T, ths, s_hist, &s_nhits, doc_base, theta, cand, cnt0, &s_emit);
        }
        if (lean) {
            // the usual case for small k: no group of the tile holds a candidate — nothing to expand, the count and the
            // threshold stay as they are (the vote is uniform, so every thread takes the same way)
            if (!__syncthreads_or(mask != 0)) {
                DI_PROF_MARK(4);
                continue;
            }
        } else {
            __syncthreads();
        }
        const uint32_t cnt_adj = cnt0 - emit_base;  // emission slots are cnt0 + (counter - its value at tile start)
        if (s_nhits - nhits_base > (uint32_t)kHistBins) {
            // Flooded: more groups than slots hold a document at or above the threshold (the first tiles of a
            // frequent-term query whose bound is still loose). In the 16-bit form every hit group still holds its
            // sums (the others are zero, i.e. below any threshold), so the accumulators can simply be re-scanned.
            if ((uint32_t)theta == 0u) {
                // No exact k-th key yet (nothing, or only a seed bound): pre-select inside the tile with a score
                // histogram. Keeping every document of the bins that hold the tile's k best is exact (at least k
                // of them score >= the cut), and the cut is a valid bound for the tiles that follow.
                int shift = 0;
                const uint32_t max_score = 255u * (uint32_t)min((uint64_t)(qe - qb), (uint64_t)65535);
                while ((max_score >> shift) >= (uint32_t)kHistBins) ++shift;
                for (uint32_t i = tid; i < (uint32_t)kHistBins; i += kScoreThreads) s_hist[i] = 0;
                __syncthreads();
                for (uint32_t g = tid; g < T / group_docs<ACC32>(); g += kScoreThreads) {
                    const uint4 x = s_acc4[g];
#pragma unroll
                    for (int i = 0; i < group_docs<ACC32>(); ++i) {
                        const uint32_t sc = group_score<ACC32>(x, i);
                        // scores above the query-length bound exist only when a posting list names a document twice:
                        // the top bin is kept whole, so clamping stays exact
                        if (sc >= ths) atomicAdd(&s_hist[min(sc >> shift, (uint32_t)kHistBins - 1u)], 1u);
                    }
                }
                __syncthreads();
                const uint32_t bin = block_find_bin_from_top(s_hist, kHistBins, p.k, s_tmp);  // 0: fewer than k documents
                if ((bin << shift) > ths) {
                    ths = bin << shift;
                    theta_pre = (uint64_t)ths << 32;
                }
            }
            __syncthreads();  // everybody has read s_nhits
            if (tid == 0) s_nhits = 0;
            nhits_base = 0;
            __syncthreads();
            scan_groups<ACC32, true>(s_acc4, T, ths, s_hist, &s_nhits, doc_base, theta, cand, cnt_adj, &s_emit);
            __syncthreads();
        }
        expand_hits<ACC32>(s_acc4, s_hist, min(s_nhits - nhits_base, (uint32_t)kHistBins), ths, doc_base, theta, cand, cnt_adj,
                           &s_emit);
        __syncthreads();
        DI_PROF_MARK(4);  // fused dense + threshold pass, emission
        uint32_t n = cnt_adj + s_emit;  // <= c0 + tile_docs <= cap
        if (lean) { emit_base = s_emit; nhits_base = s_nhits; }
        bool did_cut = false;
        dirty = dirty || n != cnt0 || theta_pre > theta;
        if (n > p.c0) {
            // too many live candidates: keep exactly the k best and raise the threshold to the k-th
            if (n <= (T * (ACC32 ? 4u : 2u)) / 8u) {
                // the accumulators are idle now: their shared memory stages the whole list (the usual case)
                theta = block_cut_to_k_staged(cand, n, p.k, reinterpret_cast<uint64_t *>(s_acc4), s_hist, s_tmp,
                                              lean ? &s_cut : &s_emit);
                if (!ACC32) zero_words16(s_acc4, (n + 1u) / 2u);  // restore the invariant (ordered by the next barrier)
                n = p.k;
            } else {
                theta = block_select_kth<true>(cand, n, p.k, s_hist, s_tmp);
                n = block_compact_ge<true>(cand, n, theta, s_scan);
            }  // theta >= theta_pre: k of the emitted keys are at or above that bound
            did_cut = true;
        } else if (theta_pre > theta) {
            theta = theta_pre;
        }
        cnt0 = n;
        // s_emit / s_hist / the accumulators are free again for the next tile (lean form: the barrier after the hit expansion
        // already says so, unless a cut just used the accumulator memory as its staging buffer)
        if ((!lean || did_cut) && sub + 1 < n_sub) __syncthreads();
        DI_PROF_MARK(5);  // cut to k
#ifdef DI_PROFILE_PHASES
        if (tid == 0 && p.prof) atomicAdd(p.prof + (size_t)tile * 8 + 7, 1ull);  // items that reached the fused pass
#endif
    }
    if (tid == 0 && dirty) {
        p.theta[sq] = theta;
        p.cnt[sq] = cnt0;
    }
    return have_state;
}

// ---- launch form A: one launch per tile, grid = queries (kept for profiling single tiles) -----
template <bool ACC32, bool BOUNDS>
__global__ void __launch_bounds__(kScoreThreads, ACC32 ? 1 : DI_SCORE_MIN_BLOCKS) score_tile_kernel(SearchArgs p, uint32_t tile)
{
    extern __shared__ uint4 s_acc4[];
    if (!ACC32) zero_words16(s_acc4, p.tile_docs / 8);  // the item's first barrier orders it
    uint32_t tma_phase = 0;
    __shared__ __align__(8) uint64_t s_mbar;  // DI_DENSE_TMA experiment only
    if (kDenseTma && threadIdx.x == 0) mbar_init(&s_mbar, 1);
    score_item<ACC32, BOUNDS>(p, tile, 1, blockIdx.x, 0, tile, tma_phase, &s_mbar);  // p.done == nullptr: the launch boundary orders the tiles
}

// ---- launch form B: ONE persistent launch for all tiles of the batch ---------------------------
// grid = resident CTAs; each CTA claims work items from a global counter in tile-major order
// (item = step * n_queries + slot, step = kTilesPerItem adjacent tiles), so at any moment the whole GPU works on a few
// tiles (their postings stay L2-resident) and there is no per-tile launch tail. After an item,
// done[q] = step + 1 is published with release semantics; the next step of the same query acquires it. Waits only
// ever point at items with a smaller index, and claims are handed out in index order, so the protocol cannot deadlock.
template <bool ACC32, bool BOUNDS>
__global__ void __launch_bounds__(kScoreThreads, ACC32 ? 1 : DI_SCORE_MIN_BLOCKS)
score_persistent_kernel(SearchArgs p, unsigned long long *counter)
{
    extern __shared__ uint4 s_acc4[];
    __shared__ unsigned long long s_item;
    const uint32_t n_virtual = p.lanes * p.n_queries;  // (lane, query) chains
    const uint32_t steps_per_lane = (p.tiles_per_lane + kTilesPerItem - 1) / kTilesPerItem;
    const unsigned long long n_items = (unsigned long long)steps_per_lane * n_virtual;
    const bool narrow = n_items <= 0xFFFFFFFFull;  // 32-bit item arithmetic (a 64-bit divide is ~100 instructions)
#ifdef DI_CLAIM_AHEAD
    unsigned long long next = 0;
    if (threadIdx.x == 0) next = atomicAdd(counter, 1ull);
#endif
    if (!ACC32) zero_words16(s_acc4, p.tile_docs / 8);  // accumulator invariant: zero at every item start
    uint32_t tma_phase = 0;
    __shared__ __align__(8) uint64_t s_mbar;  // DI_DENSE_TMA experiment only
    if (kDenseTma && threadIdx.x == 0) mbar_init(&s_mbar, 1);
#ifdef DI_CLAIM_SPLIT
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1ull);
#endif
    for (;;) {
        __syncthreads();  // everybody is done with the previous item's shared memory (DI_CLAIM_SPLIT: and the next ticket is visible)
#ifndef DI_CLAIM_SPLIT
        if (threadIdx.x == 0) {
#ifdef DI_CLAIM_AHEAD   // measured slower (profiles/README.md): the value lives across the whole item and spills
            s_item = next;
            if (next < n_items) next = atomicAdd(counter, 1ull);
#else
            s_item = atomicAdd(counter, 1ull);
#endif
        }
        __syncthreads();
#endif
        const unsigned long long item = s_item;
        if (item >= n_items) break;
        // step-major: all chains advance together, so the GPU works on `lanes` tile pairs at a time
        uint32_t step, v;
        if (narrow) {
            step = (uint32_t)item / n_virtual;
            v = (uint32_t)item - step * n_virtual;
        } else {
            step = (uint32_t)(item / n_virtual);
            v = (uint32_t)(item % n_virtual);
        }
        const uint32_t lane = p.lanes == 1 ? 0u : v / p.n_queries, slot = v - lane * p.n_queries;
        const uint32_t tile0 = lane * p.tiles_per_lane + step * kTilesPerItem;
        const uint32_t lane_end = min((lane + 1) * p.tiles_per_lane, p.n_tiles);
        if (tile0 >= lane_end) {  // the last lane may be shorter; nobody waits on these steps
#ifdef DI_CLAIM_SPLIT
            __syncthreads();  // everybody has read s_item
            if (threadIdx.x == 0) s_item = atomicAdd(counter, 1ull);
#endif
            continue;
        }
        const bool synced = score_item<ACC32, BOUNDS>(p, tile0, min((uint32_t)kTilesPerItem, lane_end - tile0), slot, lane, step, tma_phase, &s_mbar);
        // Hand-off: CTA barrier, then ONE thread publishes with a release store (MEMBAR.GPU + store). The
        // barrier orders every thread's candidate / threshold writes before the release (the pattern
        // cooperative-groups grid sync relies on). done[q] must grow one step at a time: an item that had
        // nothing to do in its tiles still waits for the previous step before announcing the next.
        __syncthreads();
#ifdef DI_CLAIM_SPLIT   // variant: the next ticket (thread 0) and the hand-off (first thread of warp 1) travel at the same time
        if (threadIdx.x == 0) s_item = atomicAdd(counter, 1ull);  // every thread read s_item before the barrier above
        if (threadIdx.x == 32) {
#else
        if (threadIdx.x == 0) {
#endif
            uint32_t *flag = p.done + lane * p.n_queries + p.recs[slot].q;
            if (!synced && step != 0) {
                uint32_t spins = 0;
                while (ld_flag_u32(flag) < step) {
                    __nanosleep(64);
                    if (++spins > (1u << 23)) __trap();
                }
            }
            st_release_u32(flag, step + 1);
        }
    }
}

}  // namespace di
