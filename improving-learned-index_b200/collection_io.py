"""Doc-major collection text -> arrays, through the library's host-side parser (csrc/collection.cu).

The reference parses "term: score, term: score" lines with Python string methods
(deep_impact_collection.py:21-25, quantize.py:21-22, 41-43, create.py:19-35). At MS MARCO scale that is
~10^9 split() calls, so the parsing (host work, no GPU involved) is done in C++ with the same
tokenisation rules, on all host threads (the file is cut into one piece per thread after a newline; the
result does not depend on the number of threads). Inputs the fast parser does not model (numeric literals
such as "1_000", bytes that are not UTF-8) fall back to `parse_python`, which is the reference's own logic
line by line.
"""
from __future__ import annotations

import ctypes
from pathlib import Path
from typing import List, Union

import numpy as np

from . import _native as N

DICT, SEQUENCE = 0, 1      # InvertedIndexCreator semantics / quantize_file semantics


class ParsedCollection:
    """doc_offsets u64[n_docs+1], term_ids u32[P] (rank of the term in sorted() order), scores f64[P], vocab (sorted)."""

    def __init__(self, doc_offsets, term_ids, scores, vocab_blob: bytes, vocab_offsets, handle=None):
        self.doc_offsets, self.term_ids, self.scores = doc_offsets, term_ids, scores
        self._vocab_blob, self._vocab_offsets = vocab_blob, vocab_offsets
        self._handle = handle

    @property
    def n_docs(self) -> int:
        return len(self.doc_offsets) - 1

    @property
    def n_terms(self) -> int:
        return len(self._vocab_offsets) - 1

    def vocab(self) -> List[str]:
        o, b = self._vocab_offsets, self._vocab_blob
        return [b[int(o[i]):int(o[i + 1])].decode('utf-8') for i in range(self.n_terms)]

    def vocab_file_bytes(self) -> bytes:
        """create.py:27-29: one term per line."""
        o, b = self._vocab_offsets, self._vocab_blob
        return b''.join(b[int(o[i]):int(o[i + 1])] + b'\n' for i in range(self.n_terms))

    def write_quantized(self, values: np.ndarray, path: Union[str, Path]):
        """quantize.py:40-47: 'term: value' for every value > 0, one line per document."""
        values = np.ascontiguousarray(values, dtype=np.int32)
        if self._handle is not None:
            N.check(N.lib().di_collection_write_quantized(self._handle, N.ptr(values), str(path).encode()))
            return
        vocab = self.vocab()
        with open(path, 'w', encoding='utf-8') as out:
            for d in range(self.n_docs):
                lo, hi = int(self.doc_offsets[d]), int(self.doc_offsets[d + 1])
                out.write(', '.join(f'{vocab[t]}: {v}' for t, v in zip(self.term_ids[lo:hi].tolist(), values[lo:hi].tolist())
                                    if v > 0) + '\n')

    def close(self):
        if self._handle is not None:
            # the numpy views die with the handle: copy nothing, just drop them
            self.doc_offsets = self.term_ids = self.scores = None
            N.lib().di_collection_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _view(ptr, n, ctype, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,))


def parse_bytes(data: bytes, mode: int) -> ParsedCollection:
    """Fast path. Raises ValueError for lines the reference would reject, NativeError(UNSUPPORTED) otherwise."""
    h = ctypes.c_void_p()
    rc = N.lib().di_collection_parse(data, len(data), mode, ctypes.byref(h))
    if rc == N.ERR_FORMAT:
        raise ValueError(N.lib().di_last_error().decode(errors='replace'))
    N.check(rc)
    n_docs, n_post, n_terms = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32()
    N.check(N.lib().di_collection_info(h, ctypes.byref(n_docs), ctypes.byref(n_post), ctypes.byref(n_terms)))
    p = [ctypes.c_void_p() for _ in range(5)]
    N.check(N.lib().di_collection_arrays(h, *[ctypes.byref(x) for x in p]))
    vocab_offsets = _view(p[4], n_terms.value + 1, ctypes.c_uint64, np.uint64).copy()
    blob = ctypes.string_at(p[3], int(vocab_offsets[-1])) if n_terms.value else b''
    return ParsedCollection(_view(p[0], n_docs.value + 1, ctypes.c_uint64, np.uint64),
                            _view(p[1], n_post.value, ctypes.c_uint32, np.uint32),
                            _view(p[2], n_post.value, ctypes.c_double, np.float64), blob, vocab_offsets, handle=h)


def parse_python(text: str, mode: int) -> ParsedCollection:
    """The reference's own loops (slow; any input CPython accepts)."""
    docs = []
    for line in _universal_lines(text):
        s = line.strip()
        if mode == DICT:
            if not s:
                docs.append([])
                continue
            d = {}
            for pair in s.split(', '):
                term, score = pair.split(': ')
                d[term] = float(score)
            docs.append(list(d.items()))
        else:
            items = []
            for t in s.split(', '):
                term, score = t.strip().split(': ')
                items.append((term, float(score)))
            docs.append(items)
    terms = sorted({t for d in docs for t, _ in d})
    tid = {t: i for i, t in enumerate(terms)}
    offs = np.zeros(len(docs) + 1, dtype=np.uint64)
    np.cumsum([len(d) for d in docs], out=offs[1:])
    ids = np.fromiter((tid[t] for d in docs for t, _ in d), dtype=np.uint32, count=int(offs[-1]))
    scores = np.fromiter((v for d in docs for _, v in d), dtype=np.float64, count=int(offs[-1]))
    enc = [t.encode('utf-8') for t in terms]
    voffs = np.zeros(len(enc) + 1, dtype=np.uint64)
    np.cumsum([len(e) for e in enc], out=voffs[1:])
    return ParsedCollection(offs, ids, scores, b''.join(enc), voffs)


def _universal_lines(text: str):
    """Lines as `for line in open(path)` yields them (universal newlines: \\n, \\r\\n, \\r)."""
    lines = text.replace('\r\n', '\n').replace('\r', '\n').split('\n')
    if lines and lines[-1] == '':
        lines.pop()                      # a trailing newline does not start another line
    return lines


def parse_file(path: Union[str, Path], mode: int) -> ParsedCollection:
    data = Path(path).read_bytes()
    try:
        return parse_bytes(data, mode)   # validates UTF-8 itself (all host threads) and defers if it is not
    except N.NativeError as e:
        if e.code != N.ERR_UNSUPPORTED:
            raise
    text = data.decode('utf-8')          # UnicodeDecodeError for invalid input, as when the reference reads the file
    return parse_python(text, mode)
