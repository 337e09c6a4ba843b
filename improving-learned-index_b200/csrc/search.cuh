// search.cuh — K4 (final exact select + sort of each query's candidates) and K5 (cross-shard
// merge). K3 and the tile-local half of K4 live in score_tile.cuh.
//
// A search = seed_theta_kernel (proven starting thresholds) -> query_order_kernel -> ONE
// score_persistent_kernel launch over all (query, tile) work items (score_tile.cuh) -> finalize_topk_kernel,
// which cuts every candidate list to the k best keys and sorts them (score descending, docid ascending)
// — replacing heapq.nlargest of inverted_index.py:62 / nano_beir_evaluator.py:128-131.
#pragma once

#include "score_tile.cuh"
#include "select.cuh"

namespace di {

constexpr uint32_t kSortSmemKeys = 4096;

// ---------------------------------------------------------------------------- K4: final select + sort
// n keys resident in shared memory (s_keys has room for smem_keys >= n entries, a power of two): cut them to the k
// best, sort descending (score desc, docid asc) and write them to out[0 .. n_out). Returns n_out; *kth receives the
// k-th key when k keys were written, else 0.
// sort_prefix (0 = everything): only the first sort_prefix keys of the row need to be in order — the row is then
// [the sort_prefix best, sorted | the rest of the top k, in any order]. That is all a shard has to deliver to the
// cross-shard merge (it reads the first k_in columns and, rarely, re-selects from the whole row), and it turns the
// 1024-key bitonic sort of a top-1000 row into one more select + a 256-key sort.
__device__ __forceinline__ uint32_t block_topk_sorted_smem(uint64_t *s_keys, uint32_t n, uint32_t k, uint32_t smem_keys,
                                                           uint64_t *out, uint32_t *s_hist, uint32_t *s_scan, uint32_t *s_tmp,
                                                           uint64_t *kth_out, uint32_t sort_prefix = 0)
{
    uint64_t *keys = s_keys;
    uint32_t k_pow2 = 1;
    while (k_pow2 < k) k_pow2 <<= 1;
    if (n > k) {
        const uint64_t kth = block_select_kth(s_keys, n, k, s_hist, s_tmp);
        if (n + k_pow2 <= smem_keys) {  // room behind the list: pack the k survivors there (any order, sorted next)
            keys = s_keys + n;
            n = block_compact_ge_unordered(s_keys, n, kth, keys, s_tmp);
        } else {
            n = block_compact_ge(s_keys, n, kth, s_scan);
        }
    }
    uint32_t n_sort = n;  // keys[0, n_sort) get sorted
    const uint32_t at = (uint32_t)(keys - s_keys);  // the survivors sit at s_keys[at, at + n): one half of the buffer must be free
    if (sort_prefix && sort_prefix < n && 2 * k_pow2 <= smem_keys && kth_out == nullptr && (at == 0 || at >= k_pow2)) {
        // partition around the sort_prefix-th key into the free half
        const uint64_t pth = block_select_kth(keys, n, sort_prefix, s_hist, s_tmp);
        uint64_t *dst = at ? s_keys : s_keys + k_pow2;
        block_partition_ge_unordered(keys, n, pth, dst, s_tmp);  // exactly sort_prefix keys in front (keys are unique)
        keys = dst;
        n_sort = sort_prefix;
    }
    uint32_t np2 = 1;
    while (np2 < n_sort) np2 <<= 1;
    // the sorted window is a power of two: it may take in a few keys of the unsorted rest (they are smaller than every
    // key of the prefix and end up behind it), or zero padding beyond the row
    for (uint32_t i = n + threadIdx.x; i < np2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(keys, np2);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = keys[i];
    if (kth_out) *kth_out = n == k ? keys[k - 1] : 0ull;
    __syncthreads();  // s_keys may be refilled by the caller
    return n;
}

// One CTA per query: cut the candidate list to the k best and sort them descending by key
// (score desc, docid asc). Sorting happens in shared memory when the list fits.
__global__ void __launch_bounds__(kScoreThreads) finalize_topk_kernel(uint64_t *cand_all, const uint32_t *cnt, uint32_t cap,
                                                                    uint32_t k, uint32_t smem_keys, uint32_t sort_prefix,
                                                                    uint64_t *out_keys, uint32_t *out_counts)
{
    extern __shared__ uint64_t s_keys[];  // smem_keys entries
    __shared__ __align__(16) uint32_t s_hist[kSelectSmemWords];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tmp[2];
    const uint32_t q = blockIdx.x;
    uint64_t *cand = cand_all + (uint64_t)q * cap;
    uint32_t n = cnt[q];
    uint64_t *out = out_keys + (uint64_t)q * k;
    if (n <= smem_keys) {  // usual case: everything happens in shared memory after one coalesced read
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = cand[i];
        __syncthreads();
        n = block_topk_sorted_smem(s_keys, n, k, smem_keys, out, s_hist, s_scan, s_tmp, nullptr, sort_prefix);
        if (threadIdx.x == 0) out_counts[q] = n;
        return;
    }
    if (n > k) {
        const uint64_t kth = block_select_kth(cand, n, k, s_hist, s_tmp);
        n = block_compact_ge(cand, n, kth, s_scan);
    }
    uint32_t np2 = 1;
    while (np2 < n) np2 <<= 1;
    if (np2 <= smem_keys) {
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

// When shards return only their k_in < k best keys, the merged top-k is exact unless some shard that
// filled its row could still hold a better key than the merged k-th: every key it did NOT return is
// below its last returned key, so the merge is proven complete iff last_returned <= merged k-th for all
// full shards (and the merge itself holds k keys). incomplete[q] = 1 marks the queries to re-run with
// full rows.
__global__ void merge_check_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ counts_in,
                                   uint32_t n_shards, uint32_t n_queries, uint32_t k_in, uint32_t k_out,
                                   const uint64_t *__restrict__ keys_out, const uint32_t *__restrict__ counts_out,
                                   uint32_t *__restrict__ incomplete)
{
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_queries; q += gridDim.x * blockDim.x) {
        const uint32_t n = counts_out[q];
        const uint64_t kth = n == k_out ? keys_out[(uint64_t)q * k_out + k_out - 1] : 0ull;  // 0: fewer than k found
        uint32_t bad = 0;
        for (uint32_t s = 0; s < n_shards; ++s) {
            const uint64_t row = (uint64_t)s * n_queries + q;
            if (counts_in[row] == k_in && keys_in[row * k_in + k_in - 1] > kth) bad = 1;
        }
        incomplete[q] = bad;
    }
}

// ---------------------------------------------------------------------------- K5 over peer memory (fused)
// The exchange step of the sharded search as ONE kernel and NO collective: rows[s] / counts[s] point at shard s's
// sorted key rows ([.][row_stride]) and counts where that shard's own search wrote them — on a peer GPU, mapped into
// this process (CUDA IPC; the loads travel over NVLink / NVSwitch). The queries are PARTITIONED over the ranks: this
// rank merges only its own slice [q_first, q_first + gridDim.x), so gathered bytes and merge work per GPU shrink with
// the number of GPUs. One CTA per query:
//   1. pull the first min(count, k_in) keys of every shard's row straight into shared memory (k_in ~ 1.25 k / G);
//   2. select the k best and sort them (same total order as one GPU: bit-identical result);
//   3. PROVE the result: a shard that holds more than k_in keys hides only keys below its k_in-th, so the merge is
//      exact iff that key <= the merged k-th;
//   4. a query that fails the proof pulls the full rows (they are already there, nothing is searched twice) and
//      selects again — in a second launch of the same shape in which the CTAs of proven queries leave at once: no host
//      round trip, no compaction of query ids, and the first pass keeps a small shared-memory footprint (several CTAs
//      per SM) instead of one sized for the rare full-row case.
constexpr uint32_t kMaxPeerShards = 64;
constexpr int kMergeThreads = 256;

// SECOND = false: first pass over every owned query with rows cut at k_in columns (small shared memory: several CTAs per
// SM); writes the result and redo[i] = 1 for a query whose proof failed. SECOND = true: the same launch shape, CTAs of
// proven queries leave at once, the others pull the full rows (shared memory sized for n_shards full rows).
template <bool SECOND>
__global__ void __launch_bounds__(kMergeThreads) merge_pull_kernel(const uint64_t *const *__restrict__ rows,
                                                                 const uint32_t *const *__restrict__ counts,
                                                                 uint32_t n_shards, uint32_t q_first, uint32_t row_stride,
                                                                 uint32_t k_in, uint32_t k, uint32_t smem_keys,
                                                                 uint64_t *__restrict__ out_keys,
                                                                 uint32_t *__restrict__ out_counts, uint32_t *__restrict__ redo,
                                                                 uint32_t *__restrict__ n_second_pass)
{
    extern __shared__ uint64_t s_keys[];  // smem_keys entries
    __shared__ __align__(16) uint32_t s_hist[kSelectSmemWords];
    __shared__ uint32_t s_scan[33];
    __shared__ uint32_t s_tmp[2];
    __shared__ uint32_t s_cnt[kMaxPeerShards], s_pref[kMaxPeerShards + 1], s_bad;
    __shared__ const uint64_t *s_row[kMaxPeerShards];
    const uint32_t i = blockIdx.x, q = q_first + i;
    if (SECOND && !redo[i]) return;
    const uint32_t lim = SECOND ? row_stride : (k_in < row_stride ? k_in : row_stride);
    if (threadIdx.x < n_shards) {  // one remote 4-byte read per shard, all in flight together
        const uint32_t c = counts[threadIdx.x][q];
        s_cnt[threadIdx.x] = c < row_stride ? c : row_stride;
        s_row[threadIdx.x] = rows[threadIdx.x] + (uint64_t)q * row_stride;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t s = 0; s < n_shards; ++s) {
            s_pref[s] = run;
            run += s_cnt[s] < lim ? s_cnt[s] : lim;
        }
        s_pref[n_shards] = run;
        s_bad = 0;
    }
    __syncthreads();
    const uint32_t total = s_pref[n_shards];
    uint32_t s = 0;
    for (uint32_t j = threadIdx.x; j < total; j += kMergeThreads) {  // flattened over the shards: loads from different peers overlap
        while (j >= s_pref[s + 1]) ++s;
        s_keys[j] = s_row[s][j - s_pref[s]];
    }
    __syncthreads();
    uint64_t kth;
    const uint32_t n_out = block_topk_sorted_smem(s_keys, total, k, smem_keys, out_keys + (uint64_t)i * k, s_hist, s_scan, s_tmp, &kth);
    if (threadIdx.x == 0) out_counts[i] = n_out;
    if (SECOND || lim == row_stride) {  // full rows: nothing is hidden
        if (!SECOND && threadIdx.x == 0) redo[i] = 0;
        return;
    }
    if (threadIdx.x < n_shards && s_cnt[threadIdx.x] > lim && s_row[threadIdx.x][lim - 1] > kth) s_bad = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        redo[i] = s_bad;
        if (s_bad && n_second_pass) atomicAdd(n_second_pass, 1u);
    }
}

// Cross-GPU barrier on the launching stream, without a collective library: flags[r] points at rank r's flag array
// ([n_ranks] words, peer-mapped). Thread s tells rank s "rank `me` reached `epoch`" with a system-scope release
// (everything this GPU wrote before — the search results — is visible to a peer that observes the flag), then waits
// until rank s has said the same to us. One CTA; epochs only grow, so the flags are never reset.
__global__ void peer_barrier_kernel(uint32_t *const *__restrict__ flags, uint32_t n_ranks, uint32_t me, uint32_t epoch)
{
    const uint32_t s = threadIdx.x;
    if (s >= n_ranks) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[s] + me), "r"(epoch) : "memory");
    const uint32_t *mine = flags[me] + s;
    uint32_t v, spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        __nanosleep(200);
        if (++spins > (1u << 24)) __trap();  // seconds: a rank that never arrives must fail loudly, not hang the box
    }
}

}  // namespace di
