"""ctypes front-end of the CPU oracle (oracle/di_oracle.c) plus a tiny pure-Python restatement.

TEST INFRASTRUCTURE ONLY — see the header of di_oracle.c. Nothing under
``improving-learned-index_b200/`` may import this module.

The C functions restate (file:line relative to /root/reference):
  * quantize            src/deep_impact/indexing/quantize.py:13-47
  * inversion           src/deep_impact/inverted_index/create.py:31-51
  * reader + scoring    src/deep_impact/inverted_index/inverted_index.py:31-62
  * in-memory twin      src/deep_impact/evaluation/nano_beir_evaluator.py:78-137
and are pinned against the reference's own outputs in tests/golden/ (oracle/make_golden.py).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> Path:
    """Compile di_oracle.c with gcc (OpenMP) into oracle/libdi_oracle.so."""
    so = _HERE / "libdi_oracle.so"
    src = _HERE / "di_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(
            ["gcc", "-O2", "-std=c99", "-fPIC", "-fopenmp", "-ffp-contract=off",
             "-shared", "-o", str(so), str(src)],
            check=True,
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        L.dio_find_max.restype = ctypes.c_double
        L.dio_find_max.argtypes = [_f64p, ctypes.c_int64]
        L.dio_scale.restype = ctypes.c_double
        L.dio_scale.argtypes = [ctypes.c_double]
        L.dio_quantize.restype = None
        L.dio_quantize.argtypes = [_f64p, ctypes.c_int64, ctypes.c_double, _i64p]
        L.dio_invert.restype = ctypes.c_int
        L.dio_invert.argtypes = [_u32p, _u8p, _u64p, ctypes.c_uint64, ctypes.c_uint32, _u64p, _u32p, _u8p]
        L.dio_serialize.restype = None
        L.dio_serialize.argtypes = [_u64p, _u32p, _u8p, ctypes.c_uint32, _u8p, _u64p]
        L.dio_term_docs.restype = ctypes.c_int64
        L.dio_term_docs.argtypes = [_u8p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, _u32p, _u8p]
        L.dio_score_topk.restype = ctypes.c_int
        L.dio_score_topk.argtypes = [
            _u8p, ctypes.c_uint64, _u64p, ctypes.c_uint32, ctypes.c_uint32, _i64p, _u64p,
            ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _u32p, _i32p, _u32p, _u64p]
        L.dio_score_topk_csr.restype = ctypes.c_int
        L.dio_score_topk_csr.argtypes = [
            _u64p, _u32p, _u8p, ctypes.c_uint32, ctypes.c_uint32, _i64p, _u64p,
            ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _u32p, _i32p, _u32p, _u64p]
        L.dio_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


# --------------------------------------------------------------------------- quantize
def find_max(scores) -> float:
    s = np.ascontiguousarray(scores, dtype=np.float64)
    return float(lib().dio_find_max(_p(s, _f64p), s.size))


def quantize(scores, max_val: float | None = None) -> np.ndarray:
    """int(score * (255 / max_val)) per element (int64, unclamped; caller drops <= 0)."""
    s = np.ascontiguousarray(scores, dtype=np.float64)
    if max_val is None:
        max_val = find_max(s)
    scale = lib().dio_scale(float(max_val))
    out = np.empty(s.size, dtype=np.int64)
    lib().dio_quantize(_p(s, _f64p), s.size, scale, _p(out, _i64p))
    return out


# --------------------------------------------------------------------------- inversion
def invert(term_ids, impacts, doc_offsets, n_terms: int):
    """Doc-major (term_ids, impacts, doc_offsets) -> (term_offsets, docids, impacts) in the
    reference's order: term asc, impact desc, docid asc."""
    t = np.ascontiguousarray(term_ids, dtype=np.uint32)
    v = np.ascontiguousarray(impacts, dtype=np.uint8)
    o = np.ascontiguousarray(doc_offsets, dtype=np.uint64)
    n_docs = o.size - 1
    toff = np.zeros(n_terms + 1, dtype=np.uint64)
    docs = np.empty(max(t.size, 1), dtype=np.uint32)
    imps = np.empty(max(t.size, 1), dtype=np.uint8)
    rc = lib().dio_invert(_p(t, _u32p), _p(v, _u8p), _p(o, _u64p), n_docs, n_terms,
                          _p(toff, _u64p), _p(docs, _u32p), _p(imps, _u8p))
    if rc != 0:
        raise RuntimeError(f"dio_invert failed: {rc}")
    return toff, docs[: t.size], imps[: t.size]


def serialize(term_offsets, docids, impacts):
    """CSR -> (dat bytes as uint8 array, idx as uint64 array of (start, end) pairs)."""
    toff = np.ascontiguousarray(term_offsets, dtype=np.uint64)
    d = np.ascontiguousarray(docids, dtype=np.uint32)
    v = np.ascontiguousarray(impacts, dtype=np.uint8)
    n_terms = toff.size - 1
    dat = np.empty(max(5 * d.size, 1), dtype=np.uint8)
    idx = np.empty(max(2 * n_terms, 1), dtype=np.uint64)
    lib().dio_serialize(_p(toff, _u64p), _p(d, _u32p), _p(v, _u8p), n_terms, _p(dat, _u8p), _p(idx, _u64p))
    return dat[: 5 * d.size], idx[: 2 * n_terms]


# --------------------------------------------------------------------------- reader / scorer
def term_docs(dat, start: int, end: int):
    dat = np.ascontiguousarray(dat, dtype=np.uint8)
    cap = max((end - start + 4) // 5, 1)
    docs = np.empty(cap, dtype=np.uint32)
    vals = np.empty(cap, dtype=np.uint8)
    n = lib().dio_term_docs(_p(dat, _u8p), dat.size, start, end, _p(docs, _u32p), _p(vals, _u8p))
    if n < 0:
        raise RuntimeError("short read")
    return docs[:n], vals[:n]


def _flatten_queries(queries):
    offs = np.zeros(len(queries) + 1, dtype=np.uint64)
    flat = []
    for i, q in enumerate(queries):
        flat.extend(int(t) for t in q)
        offs[i + 1] = len(flat)
    return np.asarray(flat, dtype=np.int64).reshape(-1), offs


def score_topk(dat, idx, n_docs: int, queries, top_k: int, tie_mode: str = "canonical", n_threads: int = 0):
    """Score lists of term ids (-1 = out-of-vocabulary) against the raw index file bytes.

    Returns (docs[Q,k] u32, scores[Q,k] i32, counts[Q] u32, postings[Q] u64)."""
    dat = np.ascontiguousarray(dat, dtype=np.uint8)
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    n_terms = idx.size // 2
    flat, offs = _flatten_queries(queries)
    nq = len(queries)
    docs = np.zeros((nq, top_k), dtype=np.uint32)
    scores = np.zeros((nq, top_k), dtype=np.int32)
    counts = np.zeros(nq, dtype=np.uint32)
    posts = np.zeros(nq, dtype=np.uint64)
    flat_c = flat if flat.size else np.zeros(1, dtype=np.int64)
    rc = lib().dio_score_topk(_p(dat, _u8p), dat.size, _p(idx, _u64p), n_terms, n_docs,
                              _p(flat_c, _i64p), _p(offs, _u64p), nq, top_k,
                              1 if tie_mode == "canonical" else 0, n_threads,
                              _p(docs, _u32p), _p(scores, _i32p), _p(counts, _u32p), _p(posts, _u64p))
    if rc != 0:
        raise RuntimeError(f"dio_score_topk failed: {rc}")
    return docs, scores, counts, posts


def score_topk_csr(term_offsets, docids, impacts, n_docs: int, queries, top_k: int,
                   tie_mode: str = "canonical", n_threads: int = 0):
    toff = np.ascontiguousarray(term_offsets, dtype=np.uint64)
    d = np.ascontiguousarray(docids, dtype=np.uint32)
    v = np.ascontiguousarray(impacts, dtype=np.uint8)
    n_terms = toff.size - 1
    flat, offs = _flatten_queries(queries)
    nq = len(queries)
    docs = np.zeros((nq, top_k), dtype=np.uint32)
    scores = np.zeros((nq, top_k), dtype=np.int32)
    counts = np.zeros(nq, dtype=np.uint32)
    posts = np.zeros(nq, dtype=np.uint64)
    flat_c = flat if flat.size else np.zeros(1, dtype=np.int64)
    d_c = d if d.size else np.zeros(1, dtype=np.uint32)
    v_c = v if v.size else np.zeros(1, dtype=np.uint8)
    rc = lib().dio_score_topk_csr(_p(toff, _u64p), _p(d_c, _u32p), _p(v_c, _u8p), n_terms, n_docs,
                                  _p(flat_c, _i64p), _p(offs, _u64p), nq, top_k,
                                  1 if tie_mode == "canonical" else 0, n_threads,
                                  _p(docs, _u32p), _p(scores, _i32p), _p(counts, _u32p), _p(posts, _u64p))
    if rc != 0:
        raise RuntimeError(f"dio_score_topk_csr failed: {rc}")
    return docs, scores, counts, posts


def max_threads() -> int:
    return int(lib().dio_max_threads())


# --------------------------------------------------------------------------- pure-Python twin
# A second, independent restatement with dicts, for tiny cases only: it follows the
# reference statement by statement and is used to cross-check the C code on CPU.
def py_quantize_line(line: str, scale: float) -> str:
    """quantize.py:41-47 for one input line."""
    data = []
    for t in line.strip().split(', '):
        term, score = t.strip().split(': ')
        val = int(float(score) * scale)
        if val > 0:
            data.append(f'{term}: {val}')
    return ', '.join(data)


def py_invert(lines):
    """create.py:19-51 on already-quantized lines -> (vocab list, dat bytes, idx bytes)."""
    import struct
    docs = []
    for line in lines:
        s = line.strip()
        docs.append({} if not s else {t: float(v) for t, v in (p.split(': ') for p in s.split(', '))})
    terms = sorted(set().union(*[d.keys() for d in docs])) if docs else []
    vocab = {t: i for i, t in enumerate(terms)}
    lists = [[] for _ in terms]
    for doc_id, item in enumerate(docs):
        for term, val in item.items():
            lists[vocab[term]].append((doc_id, int(val)))
    dat = bytearray()
    idx = bytearray()
    for postings in lists:
        start = len(dat)
        for doc_id, val in sorted(postings, key=lambda x: x[1], reverse=True):
            dat += struct.pack('I', doc_id) + struct.pack('B', val)
        idx += struct.pack('Q', start) + struct.pack('Q', len(dat))
    return terms, bytes(dat), bytes(idx)


def py_score(vocab: dict, dat: bytes, idx: bytes, query_terms, top_k: int, canonical: bool):
    """inverted_index.py:31-62 on in-memory file images."""
    import heapq
    import struct
    scores = {}
    for term in query_terms:
        tid = vocab.get(term)
        if tid is None:
            continue
        start, end = struct.unpack('QQ', idx[16 * tid: 16 * tid + 16])
        pos = start
        while pos < end:
            doc_id, value = struct.unpack('<IB', dat[pos: pos + 5])
            pos += 5
            if value == 0:
                break
            scores[doc_id] = scores.get(doc_id, 0) + value
    if canonical:
        return sorted(scores.items(), key=lambda x: (-x[1], x[0]))[:top_k]
    return heapq.nlargest(top_k, scores.items(), key=lambda x: x[1])
