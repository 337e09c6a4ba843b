"""InvertedIndex — drop-in for src/deep_impact/inverted_index/inverted_index.py:19-62.

Loads the reference's three-file index directory once into HBM (tiled layout) and
answers ``score(query_terms, top_k)`` with the CUDA kernels of csrc/search.cuh. Differences a
caller can observe, both deliberate:
  * ties are ordered by ascending docid (the reference's order depends on PYTHONHASHSEED);
  * ``score_batch`` exists, because per-query calls cannot feed a GPU.
``term_location`` / ``term_docs`` are inspection helpers and decode the file images on the host.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, List, Sequence, Tuple, Union

import numpy as np

from .. import engine
from ..utils.defaults import (DOC_SCORE_BLOCK_BYTES, INVERTED_INDEX_DATA, INVERTED_INDEX_INDEX, INVERTED_INDEX_VOCAB)

MAX_TOP_K = 65536


class BatchResults(Sequence):
    """What score_batch returns: reads like the list of per-query ``[(doc_id, score), ...]`` lists that calling the
    reference's score() in a loop gives, but holds the results as arrays (``.docids`` / ``.scores`` [n, k] and
    ``.counts`` [n]); the tuples of query i are only built when element i is asked for. 7 M tuples for the MS MARCO dev
    queries at depth 1000 cost seconds of Python; the search itself takes 30 ms."""

    def __init__(self, docids: np.ndarray, scores: np.ndarray, counts: np.ndarray):
        self.docids, self.scores, self.counts = docids, scores, counts

    def __len__(self) -> int:
        return int(self.counts.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        c = int(self.counts[i])
        return list(zip(self.docids[i, :c].tolist(), self.scores[i, :c].tolist()))

    def __eq__(self, other):
        return len(self) == len(other) and all(a == b for a, b in zip(self, other))

    def __repr__(self) -> str:
        return f"BatchResults({len(self)} queries, k={self.docids.shape[1] if self.docids.ndim == 2 else 0})"


class InvertedIndex:
    def __init__(self, index_path: Union[str, Path], doc_lo: int = 0, doc_hi: int = 0xFFFFFFFF,
                 tile_docs: int = 0, dense_ratio: int = 0, cand_slack: int = 0):
        self.index_path = Path(index_path)
        self.vocab = self._load_vocab()
        self._idx = np.fromfile(self.index_path / INVERTED_INDEX_INDEX, dtype=np.uint64)
        dat_path = self.index_path / INVERTED_INDEX_DATA
        self._dat = (np.memmap(dat_path, dtype=np.uint8, mode='r') if dat_path.stat().st_size
                     else np.zeros(0, dtype=np.uint8))
        n_terms = self._idx.size // 2
        self.device_index = engine.DeviceIndex.from_files(self._dat, self._idx[: 2 * n_terms], doc_lo, doc_hi,
                                                          tile_docs, dense_ratio, cand_slack)
        self._n_docs_hint = self.device_index.info()['max_docid_plus1']

    def _load_vocab(self):
        with open(self.index_path / INVERTED_INDEX_VOCAB, encoding='utf-8') as f:
            return {line.strip(): i for i, line in enumerate(f)}

    # ---- inspection (host) --------------------------------------------------
    def term_location(self, term):
        term_id = self.vocab.get(term, None)
        if term_id is None:
            return None, None, None
        return term_id, int(self._idx[2 * term_id]), int(self._idx[2 * term_id + 1])

    def term_docs(self, term) -> List[Tuple[int, int]]:
        term_id, start, end = self.term_location(term)
        if term_id is None or end <= start:
            return []
        n = (end - start + DOC_SCORE_BLOCK_BYTES - 1) // DOC_SCORE_BLOCK_BYTES
        raw = np.asarray(self._dat[start: start + n * DOC_SCORE_BLOCK_BYTES])
        if raw.size != n * DOC_SCORE_BLOCK_BYTES:
            raise ValueError('index data file is shorter than its .idx says')
        rec = raw.reshape(n, DOC_SCORE_BLOCK_BYTES)
        docs = rec[:, :4].copy().view('<u4').reshape(-1)
        vals = rec[:, 4]
        zeros = np.flatnonzero(vals == 0)
        stop = int(zeros[0]) if zeros.size else n          # the reader stops at the first 0 impact
        return [(int(d), int(v)) for d, v in zip(docs[:stop], vals[:stop])]

    # ---- scoring (GPU) ------------------------------------------------------
    def _term_ids(self, query_terms: Iterable[str]) -> List[int]:
        get = self.vocab.get
        return [get(t, -1) for t in query_terms]

    def _flat_term_ids(self, queries: Sequence[Iterable[str]]):
        """All queries' terms -> (flat uint32 term ids, uint64 offsets) in ONE pass over the strings (unknown term = OOV,
        inverted_index.py:43-44): the per-query lists of ids that _term_ids + engine.flatten_queries build cost more
        Python time than the GPU needs for the search."""
        from itertools import chain
        lists = [q if isinstance(q, (list, tuple, set, frozenset)) else list(q) for q in queries]
        offs = np.zeros(len(lists) + 1, dtype=np.uint64)
        np.cumsum(np.fromiter(map(len, lists), dtype=np.int64, count=len(lists)), out=offs[1:].view(np.int64))
        ids = list(map(self.vocab.get, chain.from_iterable(lists)))   # dict.get at C speed; None = not in the vocabulary
        try:
            flat = np.array(ids, dtype=np.int64)
        except TypeError:
            flat = np.array([engine.N.OOV if t is None else t for t in ids], dtype=np.int64)
        flat = flat.reshape(-1)
        return flat.astype(np.uint32), offs

    def score_batch(self, queries: Sequence[Iterable[str]], top_k: int = 1000, pinned: bool = False) -> "BatchResults":
        """score() for many queries in one GPU pass; element i is what score(queries[i]) returns."""
        if top_k <= 0 or not len(queries):
            return BatchResults(np.zeros((len(queries), 0), dtype=np.uint32), np.zeros((len(queries), 0), dtype=np.int32),
                                np.zeros(len(queries), dtype=np.uint32))
        k = min(int(top_k), max(int(self._n_docs_hint), 1))
        if k > MAX_TOP_K:
            raise ValueError(f'top_k={top_k} on {self._n_docs_hint} documents exceeds the supported {MAX_TOP_K}')
        flat, offs = self._flat_term_ids(queries)
        return BatchResults(*self.device_index.search_arrays(flat, offs, k, pinned=pinned))

    def score_id_batch(self, term_id_lists: Sequence[Sequence[int]], top_k: int = 1000, pinned: bool = False) -> "BatchResults":
        """score_batch for queries already mapped to term ids (-1 = not in the vocabulary)."""
        if top_k <= 0 or not len(term_id_lists):
            return self.score_batch([[] for _ in term_id_lists], 0)
        k = min(int(top_k), max(int(self._n_docs_hint), 1))
        if k > MAX_TOP_K:
            raise ValueError(f'top_k={top_k} on {self._n_docs_hint} documents exceeds the supported {MAX_TOP_K}')
        return BatchResults(*self.device_index.search(term_id_lists, k, pinned=pinned))

    def score(self, query_terms, top_k=1000):
        return self.score_batch([list(query_terms)], top_k)[0]
