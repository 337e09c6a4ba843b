// run_io.cu — what happens to a batch of results after the search: the run file the reference's Ranker
// writes (src/utils/datasets.py:305-324: rows "qid<TAB>pid<TAB>rank<TAB>score\n", rank from 1, append mode),
// formatted and written by all host threads straight from the result arrays, and the per-query rank facts that
// the reference's Metrics needs (src/deep_impact/evaluation/metrics.py:26-57), computed on the device from the
// result keys while they are still resident.
//
// The Python loops they replace build one tuple per hit and one f-string per row: ~7 M of each for the MS MARCO dev
// queries at depth 1000 — seconds, against a 32 ms search.
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "common.cuh"

namespace {

// decimal digits of v, two at a time from a 200-byte table, written in place (the digit count is found first)
const char kDigitPairs[201] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

inline char *put_u32(char *p, uint32_t v)
{
    const int n = v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : v < 100000 ? 5 : v < 1000000 ? 6
                  : v < 10000000 ? 7 : v < 100000000 ? 8 : v < 1000000000 ? 9 : 10;
    char *q = p + n;
    while (v >= 100) {
        const uint32_t r = v % 100;
        v /= 100;
        q -= 2;
        memcpy(q, kDigitPairs + 2 * r, 2);
    }
    if (v >= 10) memcpy(q - 2, kDigitPairs + 2 * v, 2);
    else q[-1] = (char)('0' + v);
    return p + n;
}

inline char *put_i32(char *p, int32_t v)
{
    if (v < 0) {
        *p++ = '-';
        return put_u32(p, (uint32_t)(-(int64_t)v));
    }
    return put_u32(p, (uint32_t)v);
}

unsigned io_threads(uint64_t n_rows)
{
    if (const char *e = getenv("DI_B200_IO_THREADS")) return (unsigned)std::min(256, std::max(1, atoi(e)));
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(hw, 64), n_rows >> 14));
}

}  // namespace

// A run file being appended to. di_run_writer_submit formats a batch on the calling thread (all host threads) and hands the
// formatted pieces to a background thread group that pwrite()s them at their final offsets; it returns as soon as the
// PREVIOUS batch has reached the page cache, so the formatting of batch i overlaps the copy of batch i-1 and the caller's
// arrays are free again on return. Batches reach the file in submission order.
struct di_run_writer {
    int fd = -1;
    off_t end = 0;                       // file offset of the next batch
    std::string path;
    std::vector<std::string> in_flight;  // pieces the background threads are writing
    std::thread bg;
    int bg_errno = 0;

    int wait()
    {
        if (bg.joinable()) bg.join();
        in_flight.clear();
        if (bg_errno) {
            const int e = bg_errno;
            bg_errno = 0;
            return di::set_error(DI_ERR_ARG, "write to %s failed: %s", path.c_str(), strerror(e));
        }
        return DI_OK;
    }
};

extern "C" int di_run_writer_open(const char *path, di_run_writer **out)
{
    if (!path || !out) return di::set_error(DI_ERR_ARG, "NULL argument");
    *out = nullptr;
    const int fd = open(path, O_WRONLY | O_CREAT, 0644);
    if (fd < 0) return di::set_error(DI_ERR_ARG, "cannot open %s: %s", path, strerror(errno));
    const off_t end = lseek(fd, 0, SEEK_END);  // append, like RunFile (datasets.py:310,315)
    if (end < 0) {
        close(fd);
        return di::set_error(DI_ERR_ARG, "cannot seek in %s: %s", path, strerror(errno));
    }
    di_run_writer *w = new (std::nothrow) di_run_writer();
    if (!w) {
        close(fd);
        return di::set_error(DI_ERR_NOMEM, "host allocation failed");
    }
    w->fd = fd;
    w->end = end;
    w->path = path;
    *out = w;
    return DI_OK;
}

// Query i has the id bytes qid_blob[qid_offsets[i] .. qid_offsets[i+1]) and counts[i] hits at docids[i * row_stride ...] /
// scores[i * row_stride ...] (int `{pid}` and int `{score}` exactly as Python prints them). The queries are cut into one
// contiguous piece per host thread; every piece is formatted into its own buffer and written with pwrite at its final
// offset. The file grows by exactly the bytes RunFile.writelines would have appended.
extern "C" int di_run_writer_submit(di_run_writer *w, const char *qid_blob, const uint64_t *qid_offsets, const uint32_t *docids,
                                    const int32_t *scores, const uint32_t *counts, uint32_t n_queries, uint32_t row_stride)
{
    if (!w || (n_queries && (!qid_blob || !qid_offsets || !docids || !scores || !counts)))
        return di::set_error(DI_ERR_ARG, "NULL argument");
    uint64_t n_rows = 0;
    for (uint32_t q = 0; q < n_queries; ++q) {
        if (counts[q] > row_stride) return di::set_error(DI_ERR_ARG, "counts[%u] = %u exceeds the row stride %u", q, counts[q], row_stride);
        n_rows += counts[q];
    }
    if (n_rows == 0) return DI_OK;
    const unsigned n_threads = io_threads(n_rows);
    // pieces of about equal row counts, on query boundaries
    std::vector<uint32_t> cut(n_threads + 1, n_queries);
    cut[0] = 0;
    {
        uint64_t run = 0;
        unsigned t = 1;
        for (uint32_t q = 0; q < n_queries && t < n_threads; ++q) {
            run += counts[q];
            if (run >= n_rows * t / n_threads) cut[t++] = q + 1;
        }
    }
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<std::string> bufs(n_threads);
    auto format_piece = [&](unsigned t) {
        std::string &out = bufs[t];
        uint64_t rows = 0, qid_bytes = 0;
        for (uint32_t q = cut[t]; q < cut[t + 1]; ++q) {
            rows += counts[q];
            qid_bytes += (qid_offsets[q + 1] - qid_offsets[q]) * counts[q];
        }
        out.resize(qid_bytes + rows * 36);  // pid <= 10, rank <= 10, score <= 11 digits, 3 tabs, newline
        char *p = out.data();
        for (uint32_t q = cut[t]; q < cut[t + 1]; ++q) {
            const char *qid = qid_blob + qid_offsets[q];
            const size_t qlen = (size_t)(qid_offsets[q + 1] - qid_offsets[q]);
            const uint32_t *d = docids + (size_t)q * row_stride;
            const int32_t *s = scores + (size_t)q * row_stride;
            for (uint32_t r = 0; r < counts[q]; ++r) {
                memcpy(p, qid, qlen);
                p += qlen;
                *p++ = '\t';
                p = put_u32(p, d[r]);
                *p++ = '\t';
                p = put_u32(p, r + 1);
                *p++ = '\t';
                p = put_i32(p, s[r]);
                *p++ = '\n';
            }
        }
        out.resize((size_t)(p - out.data()));
    };
    {
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(format_piece, t);
        format_piece(0);
        for (std::thread &th : pool) th.join();
    }
    const auto t_fmt = std::chrono::steady_clock::now();
    DI_TRY(w->wait());  // the previous batch is in the page cache
    std::vector<uint64_t> at(n_threads + 1, 0);
    for (unsigned t = 0; t < n_threads; ++t) at[t + 1] = at[t] + bufs[t].size();
    const off_t base = w->end;
    if (ftruncate(w->fd, base + (off_t)at[n_threads]) != 0)
        return di::set_error(DI_ERR_ARG, "cannot grow %s: %s", w->path.c_str(), strerror(errno));
    w->end = base + (off_t)at[n_threads];
    w->in_flight = std::move(bufs);
    const bool trace = getenv("DI_B200_IO_TRACE") != nullptr;
    const double fmt_ms = 1e3 * std::chrono::duration<double>(t_fmt - t_start).count();
    w->bg = std::thread([w, at, base, n_threads, trace, fmt_ms, n_rows] {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<int> rcs(n_threads, 0);
        auto write_piece = [&](unsigned t) {
            const char *p = w->in_flight[t].data();
            uint64_t left = w->in_flight[t].size(), off = (uint64_t)base + at[t];
            while (left) {
                const ssize_t n = pwrite(w->fd, p, left, (off_t)off);
                if (n < 0) {
                    if (errno == EINTR) continue;
                    rcs[t] = errno;
                    return;
                }
                p += n;
                left -= (uint64_t)n;
                off += (uint64_t)n;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(write_piece, t);
        write_piece(0);
        for (std::thread &th : pool) th.join();
        for (unsigned t = 0; t < n_threads; ++t)
            if (rcs[t]) w->bg_errno = rcs[t];
        if (trace)
            fprintf(stderr, "di_run_writer: %llu rows, %u threads, format %.1f ms, write %.1f ms (in the background)\n",
                    (unsigned long long)n_rows, n_threads, fmt_ms,
                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    });
    return DI_OK;
}

extern "C" int di_run_writer_close(di_run_writer *w)
{
    if (!w) return DI_OK;
    const int rc = w->wait();
    close(w->fd);
    delete w;
    return rc;
}

// one batch, synchronously (RunFile.write_batch): open + submit + close
extern "C" int di_write_run_file(const char *path, const char *qid_blob, const uint64_t *qid_offsets, const uint32_t *docids,
                                 const int32_t *scores, const uint32_t *counts, uint32_t n_queries, uint32_t row_stride)
{
    di_run_writer *w = nullptr;
    DI_TRY(di_run_writer_open(path, &w));
    const int rc = di_run_writer_submit(w, qid_blob, qid_offsets, docids, scores, counts, n_queries, row_stride);
    const int rc2 = di_run_writer_close(w);
    return rc != DI_OK ? rc : rc2;
}

// ---------------------------------------------------------------------------- rank facts for Metrics, on the device
namespace di {

// One warp per query: qrels of query q = sorted docids qrel_docs[qrel_offsets[q] .. qrel_offsets[q+1]). The result
// row holds keys (score << 32 | ~docid) in rank order. best_rank[q] = rank (from 1) of the first relevant hit, 0 if
// none; hits[q][j] = relevant hits with rank <= depths[j] — the integers metrics.py:31-43 derives from the run file.
__global__ void eval_ranks_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts, uint32_t n_queries,
                                  uint32_t row_stride, const uint64_t *__restrict__ qrel_offsets,
                                  const uint32_t *__restrict__ qrel_docs, const uint32_t *__restrict__ depths, uint32_t n_depths,
                                  uint32_t *__restrict__ best_rank, uint32_t *__restrict__ hits)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (q >= n_queries) return;
    const uint64_t lo = qrel_offsets[q], hi = qrel_offsets[q + 1];
    const uint32_t n = min(counts[q], row_stride);
    uint32_t best = 0xFFFFFFFFu, cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (hi > lo) {
        for (uint32_t r = lane; r < n; r += 32) {
            const uint32_t doc = key_docid(keys[(uint64_t)q * row_stride + r]);
            uint64_t a = lo, b = hi;  // binary search in the (small) sorted relevant set
            while (a < b) {
                const uint64_t m = (a + b) >> 1;
                if (qrel_docs[m] < doc) a = m + 1; else b = m;
            }
            if (a < hi && qrel_docs[a] == doc) {
                best = min(best, r + 1);
                for (uint32_t j = 0; j < n_depths; ++j) cnt[j] += (r + 1 <= depths[j]) ? 1u : 0u;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], o);
    }
    if (lane == 0) {
        best_rank[q] = best == 0xFFFFFFFFu ? 0u : best;
        for (uint32_t j = 0; j < n_depths; ++j) hits[(uint64_t)q * n_depths + j] = cnt[j];
    }
}

}  // namespace di

extern "C" int di_eval_ranks_dev(const uint64_t *d_keys, const uint32_t *d_counts, uint32_t n_queries, uint32_t row_stride,
                                 const uint64_t *d_qrel_offsets, const uint32_t *d_qrel_docs, const uint32_t *d_depths,
                                 uint32_t n_depths, uint32_t *d_best_rank, uint32_t *d_hits, void *stream)
{
    if (n_queries == 0) return DI_OK;
    if (n_depths > 8) return di::set_error(DI_ERR_ARG, "at most 8 depths per call, got %u", n_depths);
    di::eval_ranks_kernel<<<(n_queries * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        d_keys, d_counts, n_queries, row_stride, d_qrel_offsets, d_qrel_docs, d_depths, n_depths, d_best_rank, d_hits);
    DI_KERNEL_CHECK();
    return DI_OK;
}
