// common.cuh — error plumbing and small device helpers shared by every kernel file.
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/di_b200.h"

namespace di {

// ----------------------------------------------------------------------------- errors
inline char *err_buf()
{
    static thread_local char buf[512] = "";
    return buf;
}

inline int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define DI_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return di::set_error(e__ == cudaErrorMemoryAllocation ? DI_ERR_NOMEM : DI_ERR_CUDA,    \
                                 "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                                 __LINE__);                                                        \
    } while (0)

#define DI_TRY(call)                    \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != DI_OK) return rc__; \
    } while (0)

#define DI_KERNEL_CHECK() DI_CUDA(cudaGetLastError())

// RAII device buffer used for build-time scratch (freed on scope exit, also on error paths).
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t n)
    {
        release();
        bytes = n ? n : 16;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr;
            return set_error(DI_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        }
        return DI_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// RAII scratch from the device's stream-ordered memory pool: allocated and freed in stream order, so an
// asynchronous call can own scratch without a cache shared between devices, streams or threads.
struct StreamBuf {
    void *p = nullptr;
    cudaStream_t st;
    explicit StreamBuf(cudaStream_t s) : st(s) {}
    StreamBuf(const StreamBuf &) = delete;
    StreamBuf &operator=(const StreamBuf &) = delete;
    ~StreamBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t n)
    {
        static std::atomic<uint64_t> pool_tuned{0};  // keep freed blocks in the pool instead of returning them at every sync
        int dev = 0;
        cudaGetDevice(&dev);
        const uint64_t bit = 1ull << (dev & 63);
        if (!(pool_tuned.load(std::memory_order_acquire) & bit)) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_tuned.fetch_or(bit, std::memory_order_release);
        }
        cudaError_t e = cudaMallocAsync(&p, n ? n : 16, st);
        if (e != cudaSuccess) {
            p = nullptr;
            return set_error(DI_ERR_NOMEM, "cudaMallocAsync(%zu bytes) failed: %s", n, cudaGetErrorString(e));
        }
        return DI_OK;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ----------------------------------------------------------------------------- device helpers
constexpr int kWarp = 32;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 128-bit streaming load through the read-only path, not allocating in L1: postings are read
// once per (query, tile) work item; reuse across queries happens in L2.
__device__ __forceinline__ uint4 ldg_stream_v4(const uint4 *ptr)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr));
    return r;
}

// The segments of the hottest terms are read by most work items of a tile, and the six CTAs of an SM work on the
// same few tiles: letting the posting loads allocate in what is left of L1 beside the accumulators pays (measured,
// profiles/README.md). DI_DENSE_NO_L1 / DI_SPARSE_NO_L1 are the A/B switches.
__device__ __forceinline__ uint4 ldg_cached_v4(const uint4 *ptr)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
    return r;
}
__device__ __forceinline__ uint4 ldg_dense_v4(const uint4 *ptr)
{
#ifdef DI_DENSE_NO_L1
    return ldg_stream_v4(ptr);
#else
    return ldg_cached_v4(ptr);
#endif
}
__device__ __forceinline__ uint4 ldg_sparse_v4(const uint4 *ptr)
{
#ifdef DI_SPARSE_NO_L1
    return ldg_stream_v4(ptr);
#else
    return ldg_cached_v4(ptr);
#endif
}

// TMA-family bulk prefetch of a contiguous global range into L2 (fire and forget: no shared memory, no registers, no
// barrier). bytes: multiple of 16, ptr: 16-byte aligned.
__device__ __forceinline__ void prefetch_l2_bulk(const void *ptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// ---- mbarrier + TMA 1-D bulk copy global -> shared (used by the DI_DENSE_TMA experiment, score_tile.cuh)
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t *ptr)
{
    uint64_t r;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(ptr));
    return r;
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t *ptr)
{
    uint32_t r;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(ptr));
    return r;
}

// gpu-scope acquire / release on a 32-bit flag: the hand-off of a query's state between the CTAs
// that process consecutive tiles of that query inside one persistent launch
// The flag is polled with a RELAXED gpu-scope load (served by L2) on purpose: an acquire load makes
// ptxas emit CCTL.IVALL, which would wipe the SM's L1 (segment descriptors) on every work item. The
// consumer does not need L1 invalidation, because everything it then reads of the handed-off state
// is read with ld.cg (L2) and only after a CTA barrier that follows the completed flag load.
__device__ __forceinline__ uint32_t ld_flag_u32(const uint32_t *ptr)
{
    uint32_t r;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(ptr) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_u32(uint32_t *ptr, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}

// ranked-result key: larger is better. score in the high word, bit-inverted docid in the low
// word, so that equal scores order by ASCENDING docid under a descending key sort.
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t score, uint32_t docid)
{
    return (static_cast<uint64_t>(score) << 32) | static_cast<uint64_t>(~docid);
}
__host__ __device__ __forceinline__ uint32_t key_score(uint64_t k) { return static_cast<uint32_t>(k >> 32); }
__host__ __device__ __forceinline__ uint32_t key_docid(uint64_t k) { return ~static_cast<uint32_t>(k); }

inline unsigned grid_for(uint64_t n, unsigned block, unsigned max_blocks = 148u * 32u)
{
    uint64_t g = (n + block - 1) / block;
    if (g == 0) g = 1;
    return static_cast<unsigned>(g > max_blocks ? max_blocks : g);
}

}  // namespace di
