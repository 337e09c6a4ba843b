// select.cuh — block-wide exact selection primitives shared by K3/K4/K5: radix select of the k-th
// largest 64-bit key, in-place compaction, bitonic sort, and a histogram cut finder.
#pragma once

#include "common.cuh"
#include "scan_sort.cuh"

namespace di {

// ---------------------------------------------------------------------------- exact selection
// k-th largest of n unique 64-bit keys (n >= k >= 1): MSB-first radix select, 8 bits per pass.
//  * The first digit starts at the highest bit in which the keys DIFFER (one OR-reduction pass): candidate
//    scores cluster in a narrow range, so a digit aligned to the key layout would put every key into two or
//    three bins (32-way same-address shared-memory atomics) and decide nothing.
//  * As soon as the bin holding the k-th key has at most kSelectSmall keys, they are gathered and ranked
//    directly instead of running the remaining passes (usually after the first pass).
// GLOBAL_CG: keys live in global memory written by other SMs during this launch -> read through L2.
// s_hist: kSelectSmemWords 32-bit words, 8-byte aligned.
constexpr uint32_t kSelectSmall = 64;
constexpr int kSelectSmemWords = 256 + 2 * (kSelectSmall + 1) + 2;

template <bool GLOBAL_CG = false>
__device__ uint64_t block_select_kth(const uint64_t *keys, uint32_t n, uint32_t k, uint32_t *s_hist, uint32_t *s_tmp /*2*/)
{
    uint64_t *s_small = reinterpret_cast<uint64_t *>(s_hist + 256);  // [kSelectSmall] gathered keys, then the result
    uint32_t *s_ctl = s_hist + 256 + 2 * (kSelectSmall + 1);         // [0] keys in the chosen bin, [1] gather cursor
    const uint64_t first = GLOBAL_CG ? ld_cg_u64(keys) : keys[0];
    if (threadIdx.x < 2) s_tmp[threadIdx.x] = 0;
    __syncthreads();
    uint64_t diff = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) diff |= (GLOBAL_CG ? ld_cg_u64(keys + i) : keys[i]) ^ first;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, o);
    if (lane_id() == 0 && diff) {
        if ((uint32_t)diff) atomicOr(&s_tmp[0], (uint32_t)diff);
        if ((uint32_t)(diff >> 32)) atomicOr(&s_tmp[1], (uint32_t)(diff >> 32));
    }
    __syncthreads();
    diff = ((uint64_t)s_tmp[1] << 32) | s_tmp[0];
    __syncthreads();  // s_tmp is reused below
    if (diff == 0) return first;  // n == 1
    const int msb = 63 - __clzll((long long)diff);
    uint64_t mask = msb == 63 ? 0ull : ~((2ull << msb) - 1ull);  // the bits above msb are common to all keys
    uint64_t prefix = first & mask;
    uint32_t remaining = k;
    for (int shift = msb > 7 ? msb - 7 : 0;; shift = shift > 8 ? shift - 8 : 0) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = GLOBAL_CG ? ld_cg_u64(keys + i) : keys[i];
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
                        s_ctl[0] = c[j];
                        s_ctl[1] = 0;
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
        const uint32_t in_bin = s_ctl[0];
        if (shift == 0) {
            __syncthreads();
            return prefix;
        }
        if (in_bin <= kSelectSmall) {  // gather the bin's keys and rank them directly
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint64_t key = GLOBAL_CG ? ld_cg_u64(keys + i) : keys[i];
                if ((key & mask) == prefix) s_small[atomicAdd(&s_ctl[1], 1u)] = key;
            }
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < in_bin; t += blockDim.x) {
                const uint64_t mine = s_small[t];
                uint32_t above = 0;
                for (uint32_t j = 0; j < in_bin; ++j) above += s_small[j] > mine;
                if (above + 1 == remaining) s_small[kSelectSmall] = mine;  // keys are unique: exactly one writer
            }
            __syncthreads();
            const uint64_t kth = s_small[kSelectSmall];
            __syncthreads();
            return kth;
        }
        __syncthreads();
    }
}

// keeps keys >= theta, in place, order preserved; returns how many were kept
template <bool GLOBAL_CG = false>
__device__ uint32_t block_compact_ge(uint64_t *keys, uint32_t n, uint64_t theta, uint32_t *s_scan /*33*/)
{
    uint32_t out = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t key = i < n ? (GLOBAL_CG ? ld_cg_u64(keys + i) : keys[i]) : 0ull;
        const uint32_t flag = (i < n && key >= theta) ? 1u : 0u;
        uint32_t total;
        const uint32_t pos = block_exclusive_scan(flag, s_scan, total);  // barriers inside: loads are done
        if (flag) keys[out + pos] = key;                                 // out + pos <= i
        out += total;
        __syncthreads();
    }
    return out;
}

// Cuts a candidate list (global memory, n unique keys, n >= k) to its k best through a shared-memory
// staging buffer of at least n keys: one coalesced read, the radix-select passes run on shared memory,
// one write of the k survivors to keys[0..k) (unordered). Returns the k-th largest key.
__device__ __forceinline__ uint64_t block_cut_to_k_staged(uint64_t *keys, uint32_t n, uint32_t k, uint64_t *s_keys,
                                                          uint32_t *s_hist /*kSelectSmemWords*/, uint32_t *s_tmp /*2*/,
                                                          uint32_t *s_counter)
{
    for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {  // other SMs wrote them; 4 loads in flight
        uint64_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = i0 + j * blockDim.x < n ? ld_cg_u64(keys + i0 + j * blockDim.x) : 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i0 + j * blockDim.x < n) s_keys[i0 + j * blockDim.x] = v[j];
    }
    if (threadIdx.x == 0) *s_counter = 0;
    __syncthreads();
    const uint64_t kth = block_select_kth(s_keys, n, k, s_hist, s_tmp);
    const uint32_t lane = lane_id();
    for (uint32_t i0 = threadIdx.x - lane; i0 < n; i0 += blockDim.x) {  // warp-uniform; one atomic per warp
        const uint32_t i = i0 + lane;
        const uint64_t key = i < n ? s_keys[i] : 0ull;
        const bool keep = i < n && key >= kth;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        uint32_t base = 0;
        if (lane == 0 && bal) base = atomicAdd(s_counter, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) keys[base + __popc(bal & lanemask_lt())] = key;
    }
    __syncthreads();
    return kth;
}

// In-place bitonic sort, descending, n_pow2 keys (shared or global memory). Thread t of a stage handles the pair
// (i, i | j) where i is t with a zero inserted at bit log2(j): every iteration does a compare-exchange, and
// the iterations of a thread are independent (unrolled, loads first).
template <typename Ptr>
__device__ void bitonic_sort_desc(Ptr a, uint32_t n_pow2)
{
    const uint32_t half = n_pow2 >> 1;
    for (uint32_t k2 = 2; k2 <= n_pow2; k2 <<= 1) {
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
#pragma unroll 4
            for (uint32_t t = threadIdx.x; t < half; t += blockDim.x) {
                const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
                const uint32_t ixj = i | j;
                const uint64_t x = a[i], y = a[ixj];
                const bool desc_block = (i & k2) == 0;
                if (desc_block ? (x < y) : (x > y)) {
                    a[i] = y;
                    a[ixj] = x;
                }
            }
            __syncthreads();
        }
    }
}

// keys >= theta of src[0, n) -> dst[0, ...) in arbitrary order (dst must not overlap src); one shared-memory
// atomic per warp. Returns how many were kept. s_counter: one shared word.
__device__ __forceinline__ uint32_t block_compact_ge_unordered(const uint64_t *src, uint32_t n, uint64_t theta, uint64_t *dst,
                                                               uint32_t *s_counter)
{
    if (threadIdx.x == 0) *s_counter = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    for (uint32_t i0 = threadIdx.x - lane; i0 < n; i0 += blockDim.x) {  // warp-uniform
        const uint32_t i = i0 + lane;
        const uint64_t key = i < n ? src[i] : 0ull;
        const bool keep = i < n && key >= theta;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        uint32_t base = 0;
        if (lane == 0 && bal) base = atomicAdd(s_counter, (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) dst[base + __popc(bal & lanemask_lt())] = key;
    }
    __syncthreads();
    const uint32_t kept = *s_counter;
    __syncthreads();
    return kept;
}

// Two-way partition of src[0, n) into dst[0, n) (dst must not overlap src): keys >= theta packed from the front, the
// others from the back, both in arbitrary order; one shared-memory atomic per warp and side. Returns how many are >= theta.
// s_counter: two shared words.
__device__ __forceinline__ uint32_t block_partition_ge_unordered(const uint64_t *src, uint32_t n, uint64_t theta, uint64_t *dst,
                                                                 uint32_t *s_counter)
{
    if (threadIdx.x < 2) s_counter[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    for (uint32_t i0 = threadIdx.x - lane; i0 < n; i0 += blockDim.x) {  // warp-uniform
        const uint32_t i = i0 + lane;
        const uint64_t key = i < n ? src[i] : 0ull;
        const bool hi = i < n && key >= theta, lo = i < n && key < theta;
        const uint32_t bh = __ballot_sync(0xffffffffu, hi), bl = __ballot_sync(0xffffffffu, lo);
        uint32_t base_h = 0, base_l = 0;
        if (lane == 0) {
            if (bh) base_h = atomicAdd(&s_counter[0], (uint32_t)__popc(bh));
            if (bl) base_l = atomicAdd(&s_counter[1], (uint32_t)__popc(bl));
        }
        base_h = __shfl_sync(0xffffffffu, base_h, 0);
        base_l = __shfl_sync(0xffffffffu, base_l, 0);
        if (hi) dst[base_h + __popc(bh & lanemask_lt())] = key;
        if (lo) dst[n - 1 - (base_l + __popc(bl & lanemask_lt()))] = key;
    }
    __syncthreads();
    const uint32_t kept = s_counter[0];
    __syncthreads();
    return kept;
}

// largest bin b such that hist[b] + hist[b+1] + ... >= k, or 0 when the whole histogram holds fewer
// than k. n_bins is a multiple of 256. Result is returned to every thread of the block.
__device__ uint32_t block_find_bin_from_top(const uint32_t *s_hist, int n_bins, uint32_t k, uint32_t *s_tmp /*2*/)
{
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        uint32_t remaining = k;
        bool done = false;
        for (int base = n_bins - 256; base >= 0 && !done; base -= 256) {
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // lane L owns bins base+255-8L .. base+248-8L, descending
                c[j] = s_hist[base + 255 - 8 * (int)lane - j];
                sum += c[j];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (unsigned)o) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (remaining <= total) {
                const uint32_t excl = incl - sum;
                if (excl < remaining && remaining <= incl) {
                    uint32_t r = remaining - excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (r <= c[j]) {
                            s_tmp[0] = (uint32_t)(base + 255 - 8 * (int)lane - j);
                            break;
                        }
                        r -= c[j];
                    }
                }
                done = true;
            } else {
                remaining -= total;
            }
        }
        if (!done && lane == 0) s_tmp[0] = 0;
    }
    __syncthreads();
    const uint32_t bin = s_tmp[0];
    __syncthreads();
    return bin;
}

}  // namespace di
