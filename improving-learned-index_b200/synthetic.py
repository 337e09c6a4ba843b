"""Seeded synthetic collections and query sets of the shapes BASELINE.json names.

The reference ships no data and no network is available, so every test and benchmark
input is generated here: a BERT-sized vocabulary whose term popularity is
Zipf(s=1), documents made of i.i.d. Zipf draws (unique terms kept), impacts with three
decimals exactly as the reference's indexer writes them (``round(impact, 3)``,
src/deep_impact/indexing/indexer.py:62-67), and short Zipf queries.

Host-side numpy only: this is input generation, not part of the scored path. A CUDA-side
generator of the same distribution for the 8.8 M-document configuration lives in bench.py
(torch is used there as plumbing to fill device buffers).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BERT_VOCAB = 30522
MSMARCO_DOCS = 8_841_823
MSMARCO_DEV_QUERIES = 6980


def term_name(term_id: int) -> str:
    """Zero-padded so that Python's sorted(str) order equals numeric order
    (the reference assigns term ids by sorted() rank, create.py:24-25)."""
    return f"t{term_id:05d}"


def zipf_cdf(vocab_size: int, s: float = 1.0, perm_seed: int = 1234):
    """Returns (cdf over ranks, rank->term id permutation)."""
    ranks = np.arange(1, vocab_size + 1, dtype=np.float64)
    p = 1.0 / np.power(ranks, s)
    p /= p.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    perm = np.random.default_rng(perm_seed).permutation(vocab_size).astype(np.uint32)
    return cdf, perm


@dataclass
class Collection:
    """Doc-major postings: doc d owns [doc_offsets[d], doc_offsets[d+1])."""
    n_docs: int
    vocab_size: int
    doc_offsets: np.ndarray   # uint64 [n_docs+1]
    term_ids: np.ndarray      # uint32 [P]
    impacts: np.ndarray       # float64 [P]  (3-decimal values, may contain 0.0)

    def lines(self):
        """The reference's doc-major text format: 'term: score, term: score' per doc
        (deep_impact_collection.py:21-25)."""
        out = []
        for d in range(self.n_docs):
            lo, hi = int(self.doc_offsets[d]), int(self.doc_offsets[d + 1])
            out.append(', '.join(f"{term_name(int(t))}: {float(v)!r}"
                                 for t, v in zip(self.term_ids[lo:hi], self.impacts[lo:hi])))
        return out


def make_collection(n_docs: int, vocab_size: int = BERT_VOCAB, draws_per_doc: int = 120,
                    seed: int = 0, zero_frac: float = 0.01, s: float = 1.0) -> Collection:
    """n_docs documents of `draws_per_doc` Zipf draws each (duplicates removed, first
    occurrence order kept), impacts = round(1000*LogNormal(0, 0.75))/1000 clipped to
    [0, 12], with a small fraction forced to 0.0 so the `val > 0` drop path is exercised."""
    rng = np.random.default_rng(seed)
    cdf, perm = zipf_cdf(vocab_size, s)
    u = rng.random((n_docs, draws_per_doc))
    ranks = np.searchsorted(cdf, u, side='right').clip(0, vocab_size - 1)
    terms = perm[ranks]                                   # [n_docs, draws]
    # keep first occurrence of each term per row
    order = np.argsort(terms, axis=1, kind='stable')
    sorted_terms = np.take_along_axis(terms, order, axis=1)
    first = np.ones_like(sorted_terms, dtype=bool)
    first[:, 1:] = sorted_terms[:, 1:] != sorted_terms[:, :-1]
    keep = np.zeros_like(first)
    np.put_along_axis(keep, order, first, axis=1)
    counts = keep.sum(axis=1)
    doc_offsets = np.zeros(n_docs + 1, dtype=np.uint64)
    doc_offsets[1:] = np.cumsum(counts)
    term_ids = terms[keep].astype(np.uint32)
    m = np.rint(1000.0 * rng.lognormal(0.0, 0.75, size=term_ids.size)).clip(0, 12000)
    m[rng.random(term_ids.size) < zero_frac] = 0
    impacts = m / 1000.0
    return Collection(n_docs, vocab_size, doc_offsets, term_ids, impacts)


def make_queries(n_queries: int, vocab_size: int = BERT_VOCAB, seed: int = 7, mean_extra: float = 5.0,
                 max_len: int = 16, s: float = 1.0):
    """|q| = 1 + Poisson(mean_extra) clipped to [1, max_len] DISTINCT Zipf-drawn term ids."""
    rng = np.random.default_rng(seed)
    cdf, perm = zipf_cdf(vocab_size, s)
    lens = (1 + rng.poisson(mean_extra, size=n_queries)).clip(1, min(max_len, vocab_size))
    queries = []
    for n in lens:
        seen = []
        while len(seen) < n:
            r = np.searchsorted(cdf, rng.random(int(n) * 2), side='right').clip(0, vocab_size - 1)
            for t in perm[r]:
                t = int(t)
                if t not in seen:
                    seen.append(t)
                    if len(seen) == n:
                        break
        queries.append(seen)
    return queries
