// scan_sort.cuh — device-wide exclusive scan (u32) and a STABLE least-significant-digit radix
// sort of 64-bit keys over a chosen bit range. Hand-written (no CUB/Thrust on the path).
//
// The sort is what turns doc-major postings into term-major CSR (replacing the Python
// list-of-lists + sorted() of create.py:31-46) and what groups postings by (tile, term) for
// the tiled HBM layout. Payload travels inside the key, so one array is permuted per pass.
#pragma once

#include "common.cuh"

namespace di {

// ============================================================================ exclusive scan
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 elements per block

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t &total)
{
    // returns the exclusive prefix of v across the block; total = block sum. s_warp: >= 32 words.
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < nwarps ? s_warp[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += t;
        }
        s_warp[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    uint32_t res = s_warp[warp] + incl - v;
    total = s_warp[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *in, uint64_t n, uint32_t *block_sums)
{
    __shared__ uint32_t s_warp[33];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        uint64_t idx = base + (uint64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) sum += in[idx];
    }
    uint32_t total;
    block_exclusive_scan(sum, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t *in, uint32_t *out, uint64_t n, const uint32_t *block_offsets)
{
    __shared__ uint32_t s_warp[33];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        sum += v[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, s_warp, total) + (block_offsets ? block_offsets[blockIdx.x] : 0u);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

// out[i] = sum(in[0..i)); in == out allowed. `scratch` must hold scan_scratch_words(n) u32.
inline uint64_t scan_scratch_words(uint64_t n)
{
    uint64_t words = 0;
    while (n > 1) {
        n = (n + kScanTile - 1) / kScanTile;
        words += n + 1;
        if (n == 1) break;
    }
    return words + 1;
}

inline int exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *scratch, cudaStream_t st)
{
    if (n == 0) return DI_OK;
    const uint64_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 1) {
        scan_apply_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
        DI_KERNEL_CHECK();
        return DI_OK;
    }
    uint32_t *sums = scratch;
    scan_reduce_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums);
    DI_KERNEL_CHECK();
    DI_TRY(exclusive_scan_u32(sums, sums, nb, scratch + nb + 1, st));
    scan_apply_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, n, sums);
    DI_KERNEL_CHECK();
    return DI_OK;
}

// ============================================================================ stable LSD radix sort ("one sweep")
// Sorts 64-bit keys by a bit range with 8-bit digits, optionally SEGMENTED: the key array is a sequence of
// independent segments (the document tiles of a shard) that are sorted each within itself. Per call:
//   * ONE histogram kernel counts every pass's digits in a single read of the keys, one small kernel turns the
//     counts into digit bases (per segment and pass);
//   * per pass ONE kernel: a block takes a tile of 8192 keys in ticket order, ranks them (stable: warp ballots + per-warp
//     running counters), finds its offset inside every digit's output run by DECOUPLED LOOK-BACK over the preceding
//     blocks of its segment (each block first publishes its own digit counts, then sums its predecessors' until it
//     meets one that already knows its inclusive prefix), sorts the tile by digit in shared memory and writes each
//     digit's run with consecutive threads.
//   * optionally the LAST pass takes the keys apart instead of writing them (RsEpilogue below: the inversion's docid /
//     impact arrays and the start of every term come straight out of the sort).
// No separate scan kernels, no host synchronisation, no allocation inside (scratch is a caller-owned
// RadixSortScratch). Which of the two buffers holds the result is device-side state (passes whose digit is uniform are
// skipped on the device): consumers read it through rs_result().
#ifndef DI_RS_THREADS
#define DI_RS_THREADS 512
#endif
#ifndef DI_RS_ITEMS
#define DI_RS_ITEMS 16
#endif
#ifndef DI_RS_MIN_BLOCKS
#define DI_RS_MIN_BLOCKS 2
#endif
#ifndef DI_RS_MATCH_STEPS
#define DI_RS_MATCH_STEPS 0xAAAAu   // bit s set: ranking step s finds its peers with MATCH.ANY, else with ballots
#endif
constexpr int kRsThreads = DI_RS_THREADS;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = DI_RS_ITEMS;              // keys per thread (even)
static_assert(kRsThreads >= 256 && kRsThreads % 32 == 0 && kRsItems % 2 == 0 && kRsItems <= 16, "sort block shape");
constexpr int kRsTile = kRsThreads * kRsItems;     // 8192 keys per block
constexpr int kRsWarpTile = 32 * kRsItems;         // 512 contiguous keys per warp
constexpr int kRsMaxPasses = 8;
constexpr int kRsHistThreads = 256;

// look-back cell of (block, digit): 0 = nothing yet, else count + 1 in the low 30 bits and a kind in the top two
constexpr uint32_t kLbAggregate = 1u << 30, kLbPrefix = 2u << 30, kLbValue = (1u << 30) - 1u;

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Segment layout of a sort (device arrays): segment s = keys [first_key[s], first_key[s+1]) handled by the blocks
// [first_block[s], first_block[s+1]), 8192 keys each, never straddling a segment boundary.
struct SortLayout {
    const uint64_t *first_key;    // [n_segs + 1]
    const uint32_t *first_block;  // [n_segs + 1]; first_block[n_segs] = number of blocks
    uint32_t n_segs;
};

// the segment of block b (largest s with first_block[s] <= b); thread 0 only
__device__ __forceinline__ uint32_t rs_segment_of(const SortLayout &lay, uint32_t b)
{
    uint32_t lo = 0, hi = lay.n_segs;  // invariant: first_block[lo] <= b < first_block[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (lay.first_block[mid] <= b) lo = mid; else hi = mid;
    }
    return lo;
}

// layout of a plain (one segment) sort
__global__ void rs_plain_layout_kernel(uint64_t n, uint64_t *first_key, uint32_t *first_block)
{
    first_key[0] = 0;
    first_key[1] = n;
    first_block[0] = 0;
    first_block[1] = (uint32_t)((n + kRsTile - 1) / kRsTile);
}

// layout of a segmented sort from the segments' key boundaries (one block; n_segs <= 65536)
__global__ void __launch_bounds__(1024) rs_segment_blocks_kernel(const uint64_t *__restrict__ first_key, uint32_t n_segs,
                                                               uint32_t *__restrict__ first_block)
{
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (uint32_t s0 = 0; s0 < n_segs; s0 += blockDim.x) {
        const uint32_t s = s0 + threadIdx.x;
        const uint32_t nb = s < n_segs ? (uint32_t)((first_key[s + 1] - first_key[s] + kRsTile - 1) / kRsTile) : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan(nb, s_warp, total);
        const uint32_t base = s_run;
        if (s < n_segs) first_block[s] = base + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_run = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) first_block[n_segs] = s_run;
}

// hist[(seg * n_passes + p) * 256 + digit] += keys of block b with that digit in pass p, all passes in one read
__global__ void __launch_bounds__(kRsHistThreads)
rs_histogram_kernel(const uint64_t *__restrict__ keys, SortLayout lay, int bit_lo, int bit_hi, int n_passes,
                    uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[kRsMaxPasses][256];
    __shared__ uint32_t s_seg;
    const uint32_t b = blockIdx.x;
    if (b >= lay.first_block[lay.n_segs]) return;
    if (threadIdx.x == 0) s_seg = rs_segment_of(lay, b);
    for (int i = threadIdx.x; i < n_passes * 256; i += kRsHistThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t seg = s_seg;
    const uint64_t lo = lay.first_key[seg] + (uint64_t)(b - lay.first_block[seg]) * kRsTile;
    const uint64_t hi = min(lo + (uint64_t)kRsTile, lay.first_key[seg + 1]);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += kRsHistThreads) {
        const uint64_t k = keys[i];
        for (int p = 0; p < n_passes; ++p) {
            const int shift = bit_lo + 8 * p;
            const int width = bit_hi - shift < 8 ? bit_hi - shift : 8;
            atomicAdd(&h[p][(unsigned)(k >> shift) & ((1u << width) - 1u)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * 256; i += kRsHistThreads)
        if ((&h[0][0])[i]) atomicAdd(&hist[(size_t)seg * n_passes * 256 + i], (&h[0][0])[i]);
}

// per (segment, pass): counts -> absolute output position of each digit's run (in place); a pass in which some
// segment holds two different digits is marked as needed (the others are the identity and are skipped on the device)
__global__ void __launch_bounds__(256) rs_bases_kernel(uint32_t *__restrict__ hist, SortLayout lay, int n_passes,
                                                      uint32_t *__restrict__ needed)
{
    __shared__ uint32_t s_warp[33];
    const uint32_t seg = blockIdx.x / n_passes, p = blockIdx.x % n_passes;
    uint32_t *h = hist + ((size_t)seg * n_passes + p) * 256;
    const uint64_t size = lay.first_key[seg + 1] - lay.first_key[seg];
    const uint32_t c = h[threadIdx.x];
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(c, s_warp, total);
    h[threadIdx.x] = (uint32_t)lay.first_key[seg] + excl;  // n < 2^32
    if (c != 0 && c != size) needed[p] = 1u;               // benign race: everybody writes 1
}

__device__ __forceinline__ const uint64_t *rs_result(const uint64_t *a, const uint64_t *b, const uint32_t *cur)
{
    return (*cur & 1u) ? b : a;
}

// Optional epilogue of the LAST pass (the inversion): instead of writing the sorted 64-bit keys and reading them once more
// to take them apart, the pass stores the two halves of a posting where they belong — low[pos] = low 32 bits of the key,
// byte[pos] = 255 - bits 32..39 — and notes where each value of the field above `field_shift` first appears
// (first[field] = min position; every block reports the first key of each run of equal fields inside its sorted tile, and
// the first key of a field in the whole array is the first one of some tile). A pass the device skips (all keys share the
// digit) leaves this undone; the caller reads needed[last pass] to know and runs its separate pass instead.
struct RsEpilogue {
    uint32_t *low = nullptr;
    uint8_t *byte = nullptr;
    unsigned long long *first = nullptr;
    int field_shift = 0;
};

template <bool EPI>
__global__ void __launch_bounds__(kRsThreads, DI_RS_MIN_BLOCKS)
rs_onesweep_kernel(uint64_t *__restrict__ buf_a, uint64_t *__restrict__ buf_b, SortLayout lay, int shift, unsigned dmask,
                   int pass, int n_passes, const uint32_t *__restrict__ digit_base /* [n_segs][n_passes][256] */,
                   uint32_t *__restrict__ lookback /* [n_blocks][256] of this pass, zeroed */, uint32_t *__restrict__ ticket,
                   const uint32_t *__restrict__ needed_this, const uint32_t *__restrict__ cur, RsEpilogue epi)
{
    if (!*needed_this) return;  // every segment has one digit only: identity pass
    const uint64_t *__restrict__ keys_in = (*cur & 1u) ? buf_b : buf_a;
    uint64_t *__restrict__ keys_out = (*cur & 1u) ? buf_a : buf_b;

    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(rs_smem);                                    // [kRsTile]
    uint32_t(*warp_cnt)[256] = reinterpret_cast<uint32_t(*)[256]>(rs_smem + (size_t)kRsTile * 8);  // [kRsWarps][256]
    __shared__ uint32_t s_global[256];  // global start of digit d for this block, minus its local start
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_block, s_seg;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        const uint32_t b = atomicAdd(ticket, 1u);
        s_block = b;
        s_seg = b < lay.first_block[lay.n_segs] ? rs_segment_of(lay, b) : 0xFFFFFFFFu;
    }
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t blk = s_block, seg = s_seg;
    if (seg == 0xFFFFFFFFu) return;  // the grid is an upper bound of the block count
    const uint32_t seg_block0 = lay.first_block[seg];
    const uint64_t tile_lo = lay.first_key[seg] + (uint64_t)(blk - seg_block0) * kRsTile;
    const uint64_t tile_hi = min(tile_lo + (uint64_t)kRsTile, lay.first_key[seg + 1]);

    const uint64_t wbase = tile_lo + (uint64_t)warp * kRsWarpTile;
    uint64_t key[kRsItems];
    uint32_t rank2[kRsItems / 2];  // two 16-bit ranks per word (a rank is < kRsWarpTile)
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const uint64_t idx = wbase + (uint64_t)s * 32 + lane;
        key[s] = idx < tile_hi ? keys_in[idx] : 0ull;
    }
    // stable rank of every key among equal digits inside this warp's 512-key run
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const bool valid = wbase + (uint64_t)s * 32 + lane < tile_hi;
        const unsigned digit = valid ? ((unsigned)(key[s] >> shift) & dmask) : 256u;
        // Lanes holding the same digit. MATCH.ANY answers in one instruction but occupies the address-divergence unit
        // for ~50 cycles per warp (all sixteen steps on it: ADU 73 % busy, the limiter); nine ballots + masks answer on
        // the integer pipe (all sixteen there: ALU 66 % busy, the limiter). The steps alternate, so both units share
        // the work (profiles/r2_sort_ncu.txt).
        unsigned peers;
        if ((DI_RS_MATCH_STEPS >> s) & 1u) {
            peers = __match_any_sync(0xffffffffu, digit);  // invalid lanes carry digit 256: a group of their own
        } else {
            peers = __ballot_sync(0xffffffffu, valid);
            if (!valid) peers = ~peers;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const unsigned bit = (digit >> b) & 1u;
                const unsigned m = __ballot_sync(0xffffffffu, bit);
                peers &= bit ? m : ~m;
            }
        }
        uint32_t before = 0;
        if (valid) before = warp_cnt[warp][digit];
        __syncwarp();
        const uint32_t r = before + __popc(peers & lanemask_lt());
        if (valid && lane == (unsigned)(__ffs(peers) - 1)) warp_cnt[warp][digit] = before + __popc(peers);
        __syncwarp();
        if (s & 1) rank2[s >> 1] |= r << 16; else rank2[s >> 1] = r;
    }
    __syncthreads();
    uint32_t total = 0;
    const unsigned d = threadIdx.x & 255u;
    if (threadIdx.x < 256) {
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) total += warp_cnt[w][d];
        // publish this block's count of digit d, then look back for the sum of the earlier blocks of the segment
        uint32_t *cell = lookback + (size_t)blk * 256 + d;
        st_relaxed_u32(cell, (blk == seg_block0 ? kLbPrefix : kLbAggregate) | (total + 1u));
        uint32_t excl = 0;
        if (blk != seg_block0) {
            for (uint32_t t = blk - 1;; --t) {
                uint32_t v, spins = 0;
                while ((v = ld_relaxed_u32(lookback + (size_t)t * 256 + d)) == 0u) {
                    __nanosleep(32);
                    if (++spins > (1u << 24)) __trap();  // a block with a smaller ticket is always running: never expected
                }
                excl += (v & kLbValue) - 1u;
                if (v & kLbPrefix) break;
            }
            // an inclusive prefix only fits the cell below 2^30 - 1; above that the cell stays an aggregate (successors
            // simply walk further back), which keeps the sort correct for any n < 2^32
            if (excl + total < kLbValue - 1u) st_relaxed_u32(cell, kLbPrefix | (excl + total + 1u));
        }
        s_global[d] = digit_base[((size_t)seg * n_passes + pass) * 256 + d] + excl;
    }
    {   // local exclusive scan over the digits -> where each warp's run of digit d starts inside the sorted tile
        uint32_t block_total;
        uint32_t run = block_exclusive_scan(total, s_warp, block_total);  // threads >= 256 contribute 0
        if (threadIdx.x < 256) {
            s_global[d] -= run;
#pragma unroll
            for (int w = 0; w < kRsWarps; ++w) {
                const uint32_t c = warp_cnt[w][d];
                warp_cnt[w][d] = run;
                run += c;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        if (wbase + (uint64_t)s * 32 + lane < tile_hi) {
            const unsigned digit = (unsigned)(key[s] >> shift) & dmask;
            const uint32_t r = (s & 1) ? rank2[s >> 1] >> 16 : rank2[s >> 1] & 0xFFFFu;
            s_keys[warp_cnt[warp][digit] + r] = key[s];
        }
    }
    __syncthreads();
    const uint32_t n_here = (uint32_t)(tile_hi - tile_lo);
    if (!EPI) {
        for (uint32_t i = threadIdx.x; i < n_here; i += kRsThreads) {
            const uint64_t k = s_keys[i];
            keys_out[s_global[(unsigned)(k >> shift) & dmask] + i] = k;  // = global start + (i - local start)
        }
    } else {
        for (uint32_t i = threadIdx.x; i < n_here; i += kRsThreads) {
            const uint64_t k = s_keys[i];
            const uint32_t pos = s_global[(unsigned)(k >> shift) & dmask] + i;
            epi.low[pos] = (uint32_t)k;
            epi.byte[pos] = (uint8_t)(255u - ((uint32_t)(k >> 32) & 255u));
            const uint64_t field = k >> epi.field_shift;
            if (i == 0 || (s_keys[i - 1] >> epi.field_shift) != field) atomicMin(epi.first + field, (unsigned long long)pos);
        }
    }
}

// after a pass: flip the buffer state unless the pass was skipped
__global__ void rs_flip_kernel(uint32_t *cur, const uint32_t *needed_this)
{
    if (*needed_this) *cur ^= 1u;
}

// Scratch of one sort, from the stream-ordered pool (freed in stream order when it goes out of scope).
struct RadixSortScratch {
    StreamBuf hist, lookback, ctl, lay_keys, lay_blocks;
    explicit RadixSortScratch(cudaStream_t st) : hist(st), lookback(st), ctl(st), lay_keys(st), lay_blocks(st) {}
    uint32_t *cur() const { return ctl.as<uint32_t>() + 2 * kRsMaxPasses; }  // bit 0: the result is in buffer b
    const uint32_t *needed(int pass) const { return ctl.as<uint32_t>() + kRsMaxPasses + pass; }  // 0: the device skipped the pass
};

// Sorts keys ascending by bits [bit_lo, bit_hi) (stable: ties keep input order) inside every segment. `a` holds the
// input, `b` is a same-sized buffer; the result is rs_result(a, b, ws.cur()) — device-side state. d_seg_first_key:
// nullptr = one segment, else [n_segs + 1] key boundaries (device). Fully asynchronous on `st`; `ws` must outlive
// the kernels that read ws.cur().
// precounted: the caller's key-building kernel has already added every key's digits into ws.hist (zeroed by
// rs_prepare_counts below, one segment only) — the sort then does not read the keys an extra time to count them.
inline int rs_prepare_counts(RadixSortScratch &ws, int n_passes, cudaStream_t st)
{
    DI_TRY(ws.hist.alloc((size_t)n_passes * 256 * sizeof(uint32_t)));
    DI_CUDA(cudaMemsetAsync(ws.hist.p, 0, (size_t)n_passes * 256 * sizeof(uint32_t), st));
    return DI_OK;
}

// block-level digit counting for a key-building kernel: counts[pass][digit] in shared memory (zeroed by the caller),
// flushed into the global table by rs_flush_counts after a barrier
__device__ __forceinline__ void rs_count_key(uint32_t (*s_counts)[256], uint64_t key, int bit_lo, int bit_hi, int n_passes)
{
    for (int p = 0; p < n_passes; ++p) {
        const int shift = bit_lo + 8 * p;
        const int width = bit_hi - shift < 8 ? bit_hi - shift : 8;
        atomicAdd(&s_counts[p][(unsigned)(key >> shift) & ((1u << width) - 1u)], 1u);
    }
}
__device__ __forceinline__ void rs_flush_counts(uint32_t (*s_counts)[256], int n_passes, uint32_t *hist)
{
    for (int i = threadIdx.x; i < n_passes * 256; i += blockDim.x)
        if ((&s_counts[0][0])[i]) atomicAdd(&hist[i], (&s_counts[0][0])[i]);
}

inline int radix_sort_u64(uint64_t *a, uint64_t *b, uint64_t n, int bit_lo, int bit_hi, const uint64_t *d_seg_first_key,
                          uint32_t n_segs, RadixSortScratch &ws, cudaStream_t st, bool precounted = false,
                          const RsEpilogue *last_pass_epilogue = nullptr)
{
    if (n >= (1ull << 32) - 1) return set_error(DI_ERR_ARG, "radix_sort_u64: %llu keys exceed 2^32-2", (unsigned long long)n);
    const int n_passes = bit_hi > bit_lo ? (bit_hi - bit_lo + 7) / 8 : 0;
    if (n_passes > kRsMaxPasses) return set_error(DI_ERR_ARG, "radix_sort_u64: bit range too wide");
    DI_TRY(ws.ctl.alloc((2 * kRsMaxPasses + 1) * sizeof(uint32_t)));
    DI_CUDA(cudaMemsetAsync(ws.ctl.p, 0, (2 * kRsMaxPasses + 1) * sizeof(uint32_t), st));
    if (n <= 1 || n_passes == 0) return DI_OK;
    if (!d_seg_first_key) n_segs = 1;
    const uint32_t max_blocks = (uint32_t)((n + kRsTile - 1) / kRsTile) + n_segs;  // upper bound: one partial block per segment
    if (precounted && n_segs != 1) return set_error(DI_ERR_ARG, "radix_sort_u64: precounted digits need a one-segment sort");
    if (!precounted) DI_TRY(ws.hist.alloc((size_t)n_segs * n_passes * 256 * sizeof(uint32_t)));
    DI_TRY(ws.lookback.alloc((size_t)max_blocks * 256 * n_passes * sizeof(uint32_t)));
    DI_TRY(ws.lay_blocks.alloc(((size_t)n_segs + 1) * sizeof(uint32_t)));
    SortLayout lay{d_seg_first_key, ws.lay_blocks.as<uint32_t>(), n_segs};
    if (!d_seg_first_key) {
        DI_TRY(ws.lay_keys.alloc(2 * sizeof(uint64_t)));
        lay.first_key = ws.lay_keys.as<uint64_t>();
        rs_plain_layout_kernel<<<1, 1, 0, st>>>(n, ws.lay_keys.as<uint64_t>(), ws.lay_blocks.as<uint32_t>());
    } else {
        rs_segment_blocks_kernel<<<1, 1024, 0, st>>>(d_seg_first_key, n_segs, ws.lay_blocks.as<uint32_t>());
    }
    DI_KERNEL_CHECK();
    static std::atomic<uint64_t> attr_done{0};
    int dev = 0;
    DI_CUDA(cudaGetDevice(&dev));
    constexpr size_t kSmem = (size_t)kRsTile * 8 + (size_t)kRsWarps * 256 * 4;
    if (!(attr_done.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
        DI_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        DI_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    uint32_t *hist = ws.hist.as<uint32_t>();
    uint32_t *tickets = ws.ctl.as<uint32_t>(), *needed = tickets + kRsMaxPasses;
    DI_CUDA(cudaMemsetAsync(ws.lookback.p, 0, (size_t)max_blocks * 256 * n_passes * sizeof(uint32_t), st));
    if (!precounted) {
        DI_CUDA(cudaMemsetAsync(ws.hist.p, 0, (size_t)n_segs * n_passes * 256 * sizeof(uint32_t), st));
        rs_histogram_kernel<<<max_blocks, kRsHistThreads, 0, st>>>(a, lay, bit_lo, bit_hi, n_passes, hist);
        DI_KERNEL_CHECK();
    }
    rs_bases_kernel<<<n_segs * n_passes, 256, 0, st>>>(hist, lay, n_passes, needed);
    DI_KERNEL_CHECK();
    for (int p = 0; p < n_passes; ++p) {
        const int shift = bit_lo + 8 * p;
        const int width = bit_hi - shift < 8 ? bit_hi - shift : 8;
        if (last_pass_epilogue && p == n_passes - 1)
            rs_onesweep_kernel<true><<<max_blocks, kRsThreads, kSmem, st>>>(a, b, lay, shift, (1u << width) - 1u, p, n_passes, hist,
                                                                            ws.lookback.as<uint32_t>() + (size_t)p * max_blocks * 256,
                                                                            tickets + p, needed + p, ws.cur(), *last_pass_epilogue);
        else
            rs_onesweep_kernel<false><<<max_blocks, kRsThreads, kSmem, st>>>(a, b, lay, shift, (1u << width) - 1u, p, n_passes, hist,
                                                                             ws.lookback.as<uint32_t>() + (size_t)p * max_blocks * 256,
                                                                             tickets + p, needed + p, ws.cur(), RsEpilogue{});
        DI_KERNEL_CHECK();
        rs_flip_kernel<<<1, 1, 0, st>>>(ws.cur(), needed + p);
        DI_KERNEL_CHECK();
    }
    return DI_OK;
}

}  // namespace di
