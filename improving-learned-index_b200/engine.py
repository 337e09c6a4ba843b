"""Thin array-level wrappers over the C ABI: quantize / invert / serialize, and DeviceIndex,
the handle of one HBM-resident index shard. The reference-shaped classes (InvertedIndex,
InvertedIndexCreator, SparseSearch, Ranker ...) are built on these.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _native as N


# --------------------------------------------------------------------------- K1
def find_max(scores) -> float:
    """quantize.py:17-24 on the GPU: max over all scores, seeded with 0."""
    s = N.np_c(scores, np.float64).reshape(-1)
    out = ctypes.c_double(0.0)
    N.check(N.lib().di_find_max_f64(N.ptr(s), s.size, ctypes.byref(out)))
    return out.value


def quantize(scores, max_val: Optional[float] = None) -> np.ndarray:
    """quantize.py:13-14,37 on the GPU: int(score * (255 / max_val)) as int32 (caller drops <= 0)."""
    s = N.np_c(scores, np.float64).reshape(-1)
    if max_val is None:
        max_val = find_max(s)
    out = np.empty(s.size, dtype=np.int32)
    N.check(N.lib().di_quantize_f64(N.ptr(s), s.size, float(max_val), N.ptr(out)))
    return out


# --------------------------------------------------------------------------- K2
def invert(term_ids, impacts, doc_offsets, n_terms: int):
    """create.py:31-46 on the GPU: doc-major postings -> term-major CSR in the reference's
    order (term asc, impact desc, docid asc). Returns (term_offsets u64, docids u32, impacts u8)."""
    t = N.np_c(term_ids, np.uint32).reshape(-1)
    v = N.np_c(impacts, np.uint8).reshape(-1)
    o = N.np_c(doc_offsets, np.uint64).reshape(-1)
    if o.size < 1 or int(o[-1]) != t.size or v.size != t.size:
        raise ValueError("doc_offsets[-1] must equal the number of postings")
    toff = np.zeros(n_terms + 1, dtype=np.uint64)
    docs = np.empty(t.size, dtype=np.uint32)
    imps = np.empty(t.size, dtype=np.uint8)
    N.check(N.lib().di_invert(N.ptr(t), N.ptr(v), N.ptr(o), o.size - 1, n_terms, N.ptr(toff), N.ptr(docs), N.ptr(imps)))
    return toff, docs, imps


def serialize(term_offsets, docids, impacts):
    """create.py:44-51 on the GPU: CSR -> (.dat image u8[5P], .idx image u64[2V])."""
    toff = N.np_c(term_offsets, np.uint64).reshape(-1)
    d = N.np_c(docids, np.uint32).reshape(-1)
    v = N.np_c(impacts, np.uint8).reshape(-1)
    n_terms = toff.size - 1
    dat = np.empty(5 * d.size, dtype=np.uint8)
    idx = np.empty(2 * n_terms, dtype=np.uint64)
    N.check(N.lib().di_serialize(N.ptr(toff), N.ptr(d), N.ptr(v), n_terms, N.ptr(dat), N.ptr(idx)))
    return dat, idx


# --------------------------------------------------------------------------- host buffers
def pinned_empty(shape, dtype) -> np.ndarray:
    """np.empty in page-locked host memory (di_host_alloc): the result buffers of a host-level search. The memory is
    released when the last array viewing it is collected."""
    import weakref
    dtype = np.dtype(dtype)
    n = max(int(np.prod(shape)) * dtype.itemsize, 16)
    ptr = ctypes.c_void_p()
    N.check(N.lib().di_host_alloc(n, ctypes.byref(ptr)))
    buf = (ctypes.c_uint8 * n).from_address(ptr.value)
    weakref.finalize(buf, N.lib().di_host_free, ptr)       # numpy keeps `buf` alive as the base of every view
    return np.frombuffer(buf, dtype=np.uint8, count=int(np.prod(shape)) * dtype.itemsize).view(dtype).reshape(shape)


# --------------------------------------------------------------------------- queries
def flatten_queries(queries: Sequence[Iterable[int]]):
    """List of term-id lists -> (flat u32 terms, u64 offsets). Negative / None ids become OOV."""
    from itertools import chain
    lists = [q if isinstance(q, (list, tuple)) else list(q) for q in queries]
    offs = np.zeros(len(lists) + 1, dtype=np.uint64)
    if lists:
        np.cumsum(np.fromiter(map(len, lists), dtype=np.int64, count=len(lists)), out=offs[1:].view(np.int64))
    try:
        flat = np.fromiter(chain.from_iterable(lists), dtype=np.int64, count=int(offs[-1]))
    except TypeError:                      # a None among the ids
        flat = np.asarray([-1 if t is None else int(t) for t in chain.from_iterable(lists)], dtype=np.int64)
    flat[(flat < 0) | (flat > N.OOV)] = N.OOV
    return flat.astype(np.uint32).reshape(-1), offs


class DeviceIndex:
    """One index shard resident in HBM (docids in [doc_lo, doc_hi), docids stay global)."""

    def __init__(self, handle: int):
        self._h = ctypes.c_void_p(handle)

    # ---- constructors -------------------------------------------------------
    @staticmethod
    def _params(tile_docs, dense_ratio, cand_slack, flags=0):
        return N.IndexParams(tile_docs or 0, dense_ratio or 0, cand_slack or 0, flags or 0)

    @classmethod
    def from_csr(cls, term_offsets, docids, impacts, doc_lo: int = 0, doc_hi: int = N.ALL_DOCS,
                 tile_docs: int = 0, dense_ratio: int = 0, cand_slack: int = 0, flags: int = 0) -> "DeviceIndex":
        toff = N.np_c(term_offsets, np.uint64).reshape(-1)
        d = N.np_c(docids, np.uint32).reshape(-1)
        v = N.np_c(impacts, np.uint8).reshape(-1)
        if int(toff[-1]) != d.size or d.size != v.size:
            raise ValueError("term_offsets[-1] must equal len(docids) == len(impacts)")
        h = ctypes.c_void_p()
        p = cls._params(tile_docs, dense_ratio, cand_slack, flags)
        N.check(N.lib().di_index_create_csr(N.ptr(toff), N.ptr(d), N.ptr(v), toff.size - 1, doc_lo, doc_hi,
                                            ctypes.byref(p), ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def from_csr_device(cls, d_term_offsets, d_docids, d_impacts, n_terms: int, n_postings: int,
                        doc_lo: int = 0, doc_hi: int = N.ALL_DOCS, tile_docs: int = 0, dense_ratio: int = 0,
                        cand_slack: int = 0, flags: int = 0) -> "DeviceIndex":
        """CSR already in device memory (torch CUDA tensors or raw device addresses). The caller
        must have synchronised the stream that produced them."""
        h = ctypes.c_void_p()
        p = cls._params(tile_docs, dense_ratio, cand_slack, flags)
        N.check(N.lib().di_index_create_csr_dev(N.ptr(d_term_offsets), N.ptr(d_docids), N.ptr(d_impacts), n_terms,
                                                n_postings, doc_lo, doc_hi, ctypes.byref(p), ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def from_docmajor_device(cls, d_term_ids, d_impacts, d_doc_offsets, n_docs: int, n_terms: int, n_postings: int,
                             doc_lo: int = 0, tile_docs: int = 0, dense_ratio: int = 0, cand_slack: int = 0,
                             flags: int = 0) -> "DeviceIndex":
        """Straight from a doc-major collection in device memory (u32 term ids, u8 impacts, u64 doc offsets of
        documents doc_lo .. doc_lo + n_docs): no term-major detour, one segmented two-pass sort. Identical index."""
        h = ctypes.c_void_p()
        p = cls._params(tile_docs, dense_ratio, cand_slack, flags)
        N.check(N.lib().di_index_create_docmajor_dev(N.ptr(d_term_ids), N.ptr(d_impacts), N.ptr(d_doc_offsets), n_docs,
                                                     n_terms, n_postings, doc_lo, ctypes.byref(p), ctypes.byref(h)))
        return cls(h.value)

    @classmethod
    def from_files(cls, dat, idx_pairs, doc_lo: int = 0, doc_hi: int = N.ALL_DOCS, tile_docs: int = 0,
                   dense_ratio: int = 0, cand_slack: int = 0, flags: int = 0) -> "DeviceIndex":
        """From the reference's file images: inverted_index.dat bytes and inverted_index.idx as u64 pairs."""
        dat = N.np_c(dat, np.uint8).reshape(-1)
        idx = N.np_c(idx_pairs, np.uint64).reshape(-1)
        if idx.size % 2:
            raise ValueError(".idx image must hold (start, end) pairs")
        h = ctypes.c_void_p()
        p = cls._params(tile_docs, dense_ratio, cand_slack, flags)
        N.check(N.lib().di_index_create_files(N.ptr(dat), dat.size, N.ptr(idx), idx.size // 2, doc_lo, doc_hi,
                                              ctypes.byref(p), ctypes.byref(h)))
        return cls(h.value)

    # ---- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().di_index_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __reduce__(self):
        raise TypeError("a DeviceIndex lives in GPU memory and cannot be pickled (the reference's "
                        "multiprocessing.Pool fan-out, ranker.py:44-46, is replaced by batched search)")

    # ---- queries ------------------------------------------------------------
    def info(self) -> dict:
        i = N.IndexInfo()
        N.check(N.lib().di_index_get_info(self._h, ctypes.byref(i)))
        return {f: getattr(i, f) for f, _ in N.IndexInfo._fields_}

    def term_df(self, term_ids) -> np.ndarray:
        t = N.np_c(term_ids, np.uint32).reshape(-1)
        out = np.zeros(t.size, dtype=np.uint64)
        N.check(N.lib().di_index_term_df(self._h, N.ptr(t), t.size, N.ptr(out)))
        return out

    def search_flat(self, q_terms, q_offsets, top_k: int, out_docids=None, out_scores=None, out_counts=None):
        """Host-buffer entry point (H2D + kernels + D2H inside the call). Buffers may be numpy arrays
        or pinned torch tensors; outputs are allocated when not given."""
        n_q = len(q_offsets) - 1
        if out_docids is None:
            out_docids = np.empty((n_q, top_k), dtype=np.uint32)
            out_scores = np.empty((n_q, top_k), dtype=np.int32)
            out_counts = np.zeros(n_q, dtype=np.uint32)
        N.check(N.lib().di_search(self._h, N.ptr(q_terms), N.ptr(q_offsets), n_q, top_k,
                                  N.ptr(out_docids), N.ptr(out_scores), N.ptr(out_counts)))
        return out_docids, out_scores, out_counts

    def search(self, queries: Sequence[Iterable[int]], top_k: int, pinned: bool = False):
        """queries: list of term-id lists. Returns (docids[Q,k], scores[Q,k], counts[Q]). pinned=True: the arrays
        live in page-locked memory owned by the index (full-rate device-to-host copy of large results) and are
        RECYCLED: two sets per result shape alternate, so a result stays valid until the second next pinned call."""
        return self.search_arrays(*flatten_queries(queries), top_k, pinned=pinned)

    def search_arrays(self, flat: np.ndarray, offs: np.ndarray, top_k: int, pinned: bool = False):
        """search() for queries already flattened: flat uint32 term ids (OOV = 0xFFFFFFFF), uint64 offsets [n + 1]."""
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint32)
        if not pinned:
            return self.search_flat(flat, offs, top_k)
        n_q = len(offs) - 1
        sets = self.__dict__.setdefault("_pinned_sets", {})
        if len(sets) > 8:
            sets.clear()
        ring = sets.setdefault((n_q, top_k), [None, None, 0])
        slot = ring[2] % 2
        ring[2] += 1
        if ring[slot] is None:
            ring[slot] = (pinned_empty((n_q, top_k), np.uint32), pinned_empty((n_q, top_k), np.int32), pinned_empty((n_q,), np.uint32))
        out = ring[slot]
        out[2][:] = 0
        return self.search_flat(flat, offs, top_k, *out)

    def search_device(self, d_q_terms, d_q_offsets, n_queries: int, max_query_len: int, top_k: int,
                      d_out_keys, d_out_counts, stream: int = 0, d_theta_init=None):
        """Device-buffer entry point, asynchronous on `stream`; outputs are packed keys
        (score << 32 | ~docid) sorted descending — the form the cross-shard merge consumes.
        d_theta_init: optional per-query keys known to be lower bounds of the final k-th best key."""
        N.check(N.lib().di_search_dev(self._h, N.ptr(d_q_terms), N.ptr(d_q_offsets), n_queries, max_query_len, top_k,
                                      N.ptr(d_theta_init), N.ptr(d_out_keys), N.ptr(d_out_counts), stream))

    def export_seed_hist(self, d_hist, stream: int = 0):
        """This shard's impact histogram of every term into d_hist [n_terms, 256] (uint32 on the device)."""
        N.check(N.lib().di_index_export_seed_hist_dev(self._h, N.ptr(d_hist), stream))

    def import_seed_hist(self, d_hist, stream: int = 0):
        """New seed tables from a histogram summed over all shards: searches then start from a bound of the GLOBAL k-th
        score and return only what can be in the top-k of the whole collection (rows for the cross-shard merge)."""
        N.check(N.lib().di_index_import_seed_hist_dev(self._h, N.ptr(d_hist), stream))

    def set_sorted_prefix(self, p: int):
        """Row order of search_device results from now on: 0 = fully sorted; p > 0 = [the p best keys, sorted | the rest
        of the top-k in any order] — what a shard owes the cross-shard merge (di_index_set_sorted_prefix)."""
        N.check(N.lib().di_index_set_sorted_prefix(self._h, int(p)))

    def timings(self) -> dict:
        t = N.Timings()
        N.check(N.lib().di_get_timings(self._h, ctypes.byref(t)))
        return {f: getattr(t, f) for f, _ in N.Timings._fields_}


def unpack_keys_device(d_keys, n: int, d_docids, d_scores, stream: int = 0):
    N.check(N.lib().di_unpack_keys_dev(N.ptr(d_keys), n, N.ptr(d_docids), N.ptr(d_scores), stream))


def merge_topk_device(d_keys_in, d_counts_in, n_shards: int, n_queries: int, top_k: int, d_keys_out, d_counts_out,
                      stream: int = 0, k_in: int = 0, d_incomplete=None):
    """K5: [n_shards][n_queries][k_in] gathered keys -> global top-k per query. With k_in < top_k,
    d_incomplete[q] = 1 marks queries whose merge could not be proven exact (re-run them with full rows)."""
    N.check(N.lib().di_merge_topk_dev(N.ptr(d_keys_in), N.ptr(d_counts_in), n_shards, n_queries, k_in or top_k, top_k,
                                      N.ptr(d_keys_out), N.ptr(d_counts_out), N.ptr(d_incomplete), stream))


def merge_pull_device(d_row_ptrs, d_count_ptrs, n_shards: int, q_first: int, n_queries: int, row_stride: int, k_in: int,
                      top_k: int, d_keys_out, d_counts_out, stream: int = 0, d_n_second_pass=None):
    """K5 fused with its exchange: d_row_ptrs / d_count_ptrs are device tables of n_shards pointers to the shards' own
    sorted rows (peer memory); merges queries [q_first, q_first + n_queries) — pull, select, prove, second pass —
    in one kernel."""
    N.check(N.lib().di_merge_pull_dev(N.ptr(d_row_ptrs), N.ptr(d_count_ptrs), n_shards, q_first, n_queries, row_stride, k_in,
                                      top_k, N.ptr(d_keys_out), N.ptr(d_counts_out), N.ptr(d_n_second_pass), stream))


def peer_barrier_device(d_flag_ptrs, n_ranks: int, my_rank: int, epoch: int, stream: int = 0):
    N.check(N.lib().di_peer_barrier_dev(N.ptr(d_flag_ptrs), n_ranks, my_rank, epoch, stream))
