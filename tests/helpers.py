"""Helpers shared by the GPU parity tests."""
import numpy as np

from improving_learned_index_b200 import synthetic as syn
from oracle import oracle


def write_index_dir(path, vocab, idx_bytes, dat_bytes):
    path.mkdir(parents=True, exist_ok=True)
    (path / "vocab.txt").write_text(''.join(t + '\n' for t in vocab), encoding='utf-8')
    (path / "inverted_index.idx").write_bytes(idx_bytes)
    (path / "inverted_index.dat").write_bytes(dat_bytes)
    return path


def canonical(pairs, k):
    """SURVEY.md §8a definition 2: the reference's FULL list re-sorted by (-score, docid), cut at k."""
    return [list(p) for p in sorted(pairs, key=lambda x: (-x[1], x[0]))[:k]]


def quantized_csr(n_docs, vocab_size, draws, seed, zero_frac=0.01):
    """Synthetic collection -> oracle-quantized doc-major arrays + oracle CSR."""
    c = syn.make_collection(n_docs, vocab_size=vocab_size, draws_per_doc=draws, seed=seed, zero_frac=zero_frac)
    q = oracle.quantize(c.impacts)
    keep = q > 0
    doc_of = np.repeat(np.arange(c.n_docs), np.diff(c.doc_offsets.astype(np.int64)))
    offs = np.zeros(c.n_docs + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(np.bincount(doc_of[keep], minlength=c.n_docs))
    terms, imps = c.term_ids[keep], q[keep].astype(np.uint8)
    toff, docs, vals = oracle.invert(terms, imps, offs, vocab_size)
    return dict(collection=c, terms=terms, imps=imps, offs=offs, toff=toff, docs=docs, vals=vals,
                n_docs=n_docs, vocab_size=vocab_size)


def assert_same_results(got, want, label=""):
    gd, gs, gc = got
    wd, ws, wc = want[:3]
    assert np.array_equal(gc, wc), f"{label}: counts differ at {np.flatnonzero(gc != wc)[:5]}"
    for i in range(len(gc)):
        n = int(gc[i])
        assert np.array_equal(gs[i, :n], ws[i, :n]), f"{label}: scores differ for query {i}"
        assert np.array_equal(gd[i, :n], wd[i, :n]), f"{label}: docids differ for query {i}"
