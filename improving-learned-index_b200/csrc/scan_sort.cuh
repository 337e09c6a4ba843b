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

// ============================================================================ stable LSD radix sort
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                       // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;     // 4096 keys per block
constexpr int kRsWarpTile = 32 * kRsItems;         // 512 contiguous keys per warp

__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, int shift, unsigned dmask,
               uint32_t *__restrict__ block_hist, uint32_t nblocks)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll 4
    for (int i = 0; i < kRsItems; ++i) {
        uint64_t idx = base + (uint64_t)i * kRsThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(unsigned)(keys[idx] >> shift) & dmask], 1u);
    }
    __syncthreads();
    block_hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// totals[d] = number of keys with digit d (used on the host to skip passes whose digit is uniform)
__global__ void rs_digit_totals_kernel(const uint32_t *block_hist_scanned, uint32_t nblocks, uint64_t n,
                                       uint32_t *totals)
{
    // exclusive-scanned digit-major table: start of digit d = scanned[d*nblocks]; end = start of d+1 (or n)
    unsigned d = threadIdx.x;
    uint64_t start = block_hist_scanned[(uint64_t)d * nblocks];
    uint64_t end = d == 255 ? n : block_hist_scanned[(uint64_t)(d + 1) * nblocks];
    totals[d] = (uint32_t)(end - start);
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint64_t *__restrict__ keys_in, uint64_t *__restrict__ keys_out, uint64_t n, int shift,
                  unsigned dmask, const uint32_t *__restrict__ block_offsets /* scanned, digit-major */, uint32_t nblocks)
{
    __shared__ uint32_t warp_cnt[kRsWarps][256];
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();

    const uint64_t wbase = (uint64_t)blockIdx.x * kRsTile + (uint64_t)warp * kRsWarpTile;
    uint64_t key[kRsItems];
    uint32_t rank[kRsItems];
    // pass 1: stable rank of every key among equal digits inside this warp's 512-key run
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const uint64_t idx = wbase + (uint64_t)s * 32 + lane;
        const bool valid = idx < n;
        key[s] = valid ? keys_in[idx] : 0ull;
        const unsigned digit = valid ? ((unsigned)(key[s] >> shift) & dmask) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        uint32_t before = 0;
        if (valid) before = warp_cnt[warp][digit];
        __syncwarp();
        rank[s] = before + __popc(peers & lanemask_lt());
        if (valid && lane == (unsigned)(__ffs(peers) - 1)) warp_cnt[warp][digit] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // Block-local sort first: keys are placed in shared memory grouped by digit (stable), so that the
    // global scatter below writes each digit's run with consecutive threads -> consecutive addresses.
    // A direct scatter from registers would touch up to 32 different 32-byte sectors per warp store.
    __shared__ uint64_t s_keys[kRsTile];
    __shared__ uint32_t s_global[256];       // global start of digit d for this block, minus s_digit_start[d]
    __shared__ uint32_t s_warp[33];
    {
        const unsigned d = threadIdx.x;      // blockDim.x == 256 == number of digits
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) total += warp_cnt[w][d];
        uint32_t block_total;
        uint32_t run = block_exclusive_scan(total, s_warp, block_total);
        s_global[d] = block_offsets[(uint64_t)d * nblocks + blockIdx.x] - run;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kRsItems; ++s) {
        const uint64_t idx = wbase + (uint64_t)s * 32 + lane;
        if (idx < n) {
            const unsigned digit = (unsigned)(key[s] >> shift) & dmask;
            s_keys[warp_cnt[warp][digit] + rank[s]] = key[s];
        }
    }
    __syncthreads();
    const uint64_t block_base = (uint64_t)blockIdx.x * kRsTile;
    const uint32_t n_here = (uint32_t)(n - block_base < (uint64_t)kRsTile ? n - block_base : (uint64_t)kRsTile);
    for (uint32_t i = threadIdx.x; i < n_here; i += kRsThreads) {
        const uint64_t k = s_keys[i];
        keys_out[s_global[(unsigned)(k >> shift) & dmask] + i] = k;  // = global start + (i - local start)
    }
}

struct RadixSortScratch {
    DevBuf hist;     // 256 * nblocks u32
    DevBuf scan;     // scan scratch
    DevBuf totals;   // 256 u32
    uint32_t nblocks = 0;
    int prepare(uint64_t n)
    {
        nblocks = (uint32_t)((n + kRsTile - 1) / kRsTile);
        if (nblocks == 0) nblocks = 1;
        DI_TRY(hist.alloc((size_t)256 * nblocks * sizeof(uint32_t)));
        DI_TRY(scan.alloc((size_t)scan_scratch_words((uint64_t)256 * nblocks) * sizeof(uint32_t)));
        DI_TRY(totals.alloc(256 * sizeof(uint32_t)));
        return DI_OK;
    }
};

// Sorts keys ascending by bits [bit_lo, bit_hi) (stable: ties keep input order). `a` holds the
// input; `b` is a same-sized buffer. *result points at whichever buffer holds the output.
// Synchronises the stream once per pass (reads 256 digit totals to skip uniform digits).
inline int radix_sort_u64(uint64_t *a, uint64_t *b, uint64_t n, int bit_lo, int bit_hi, RadixSortScratch &ws,
                          cudaStream_t st, uint64_t **result)
{
    *result = a;
    if (n >= (1ull << 32)) return set_error(DI_ERR_ARG, "radix_sort_u64: %llu keys exceed 2^32-1", (unsigned long long)n);
    if (n <= 1 || bit_hi <= bit_lo) return DI_OK;
    DI_TRY(ws.prepare(n));
    uint32_t *hist = ws.hist.as<uint32_t>();
    const uint32_t nb = ws.nblocks;
    uint64_t *src = a, *dst = b;
    for (int shift = bit_lo; shift < bit_hi; shift += 8) {
        const int width = bit_hi - shift < 8 ? bit_hi - shift : 8;
        const unsigned dmask = (1u << width) - 1u;  // never look above bit_hi
        rs_hist_kernel<<<nb, kRsThreads, 0, st>>>(src, n, shift, dmask, hist, nb);
        DI_KERNEL_CHECK();
        DI_TRY(exclusive_scan_u32(hist, hist, (uint64_t)256 * nb, ws.scan.as<uint32_t>(), st));
        rs_digit_totals_kernel<<<1, 256, 0, st>>>(hist, nb, n, ws.totals.as<uint32_t>());
        DI_KERNEL_CHECK();
        uint32_t totals[256];
        DI_CUDA(cudaMemcpyAsync(totals, ws.totals.p, sizeof totals, cudaMemcpyDeviceToHost, st));
        DI_CUDA(cudaStreamSynchronize(st));
        bool uniform = false;
        for (int d = 0; d < 256; ++d)
            if (totals[d] == n) uniform = true;
        if (uniform) continue;  // every key has the same digit: the pass is the identity
        rs_scatter_kernel<<<nb, kRsThreads, 0, st>>>(src, dst, n, shift, dmask, hist, nb);
        DI_KERNEL_CHECK();
        uint64_t *t = src;
        src = dst;
        dst = t;
    }
    *result = src;
    return DI_OK;
}

}  // namespace di
