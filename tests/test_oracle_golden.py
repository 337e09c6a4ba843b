"""Pins the CPU oracle (oracle/di_oracle.c + the pure-Python twin) against outputs of the
REFERENCE'S OWN CODE recorded in tests/golden/ by oracle/make_golden.py."""
import hashlib

import numpy as np
import pytest

from oracle import oracle
import improving_learned_index_b200.synthetic as syn


def parse_quantized(lines, vocab):
    """Already-quantized doc-major lines -> (term_ids, impacts, doc_offsets) with the
    reference's dict semantics (deep_impact_collection.py:21-25: last duplicate wins)."""
    tid = {t: i for i, t in enumerate(vocab)}
    terms, imps, offs = [], [], [0]
    for line in lines:
        s = line.strip()
        d = {} if not s else {t: int(float(v)) for t, v in (p.split(': ') for p in s.split(', '))}
        for t, v in d.items():
            terms.append(tid[t])
            imps.append(v)
        offs.append(len(terms))
    return (np.asarray(terms, dtype=np.uint32), np.asarray(imps, dtype=np.uint8),
            np.asarray(offs, dtype=np.uint64))


def as_pairs(docs, scores, count):
    return [[int(d), int(s)] for d, s in zip(docs[:count], scores[:count])]


# ------------------------------------------------------------------ quantize
def test_quantize_cases(golden):
    g = golden("quantize")
    for case in g["cases"]:
        got = oracle.quantize(case["values"], case["max"])
        assert got.tolist() == case["quantized"], case["max"]


def test_quantize_self_max_sweep(golden):
    g = golden("quantize")
    vals = np.arange(1, 20001) / 1000
    got = np.array([oracle.quantize([v], v)[0] for v in vals])
    assert (np.nonzero(got == 254)[0] + 1).tolist() == g["sweep_254"]
    assert g["sweep_other"] == [] and set(got.tolist()) <= {254, 255}
    assert len(g["sweep_254"]) == 2722          # the count SURVEY.md §7 reports


def test_quantize_file_lines(golden):
    g = golden("quantize")["file"]
    scores = [float(p.split(': ')[1]) for l in g["lines"] for p in l.split(', ')]
    for key, mx in (("auto", None), ("max2", 2.0), ("max05", 0.5)):
        m = oracle.find_max(scores) if mx is None else mx
        scale = 255 / m
        assert [oracle.py_quantize_line(l, scale) for l in g["lines"]] == g[key]
        # C path on the same numbers
        q = oracle.quantize(scores, m).tolist()
        it = iter(q)
        rebuilt = []
        for l in g["lines"]:
            parts = []
            for p in l.split(', '):
                v = next(it)
                if v > 0:
                    parts.append(f"{p.split(': ')[0]}: {v}")
            rebuilt.append(', '.join(parts))
        assert rebuilt == g[key]


# ------------------------------------------------------------------ inversion + file format
@pytest.mark.parametrize("name", ["kat", "small", "zeros"])
def test_invert_bytes_identical(golden, name):
    g = golden(name)
    lines = g["quantized_lines"] if "quantized_lines" in g else g["lines"]
    vocab, dat_py, idx_py = oracle.py_invert(lines)
    assert vocab == g["vocab"]
    assert dat_py == g["dat"] and idx_py == g["idx"]
    t, v, o = parse_quantized(lines, g["vocab"])
    toff, docs, imps = oracle.invert(t, v, o, len(g["vocab"]))
    dat, idx = oracle.serialize(toff, docs, imps)
    assert dat.tobytes() == g["dat"]
    assert idx.tobytes() == g["idx"]


def test_medium_collection_reproduces(golden):
    """The committed generator + oracle regenerate the 3000-doc reference build bit for bit."""
    g = golden("medium")
    c = syn.make_collection(**{k: g["gen"][k] for k in ("n_docs", "vocab_size", "draws_per_doc", "seed")})
    lines = c.lines()
    sha = lambda b: hashlib.sha256(b).hexdigest()
    assert sha(''.join(l + '\n' for l in lines).encode()) == g["lines_sha256"]
    q = oracle.quantize(c.impacts)
    keep = q > 0
    used = sorted(set(c.term_ids[keep].tolist()))
    vocab = [syn.term_name(t) for t in used]
    assert sha(''.join(t + '\n' for t in vocab).encode()) == g["vocab_sha256"]
    remap = np.full(c.vocab_size, -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    doc_of = np.repeat(np.arange(c.n_docs), np.diff(c.doc_offsets.astype(np.int64)))
    offs = np.zeros(c.n_docs + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(np.bincount(doc_of[keep], minlength=c.n_docs))
    toff, docs, imps = oracle.invert(remap[c.term_ids[keep]], q[keep], offs, len(used))
    dat, idx = oracle.serialize(toff, docs, imps)
    assert dat.size // 5 == g["n_postings"]
    assert sha(dat.tobytes()) == g["dat_sha256"]
    assert sha(idx.tobytes()) == g["idx_sha256"]
    # scoring: raw reference order (ties in first-touch order) for 30 queries, k=1000
    tid = {t: i for i, t in enumerate(vocab)}
    queries = [[tid.get(t, -1) for t in q["terms"]] for q in g["queries"]]
    d, s, cnt, _ = oracle.score_topk(dat, idx, c.n_docs, queries, 1000, tie_mode="raw")
    for i, q in enumerate(g["queries"]):
        assert as_pairs(d[i], s[i], cnt[i]) == q["top1000"], i
    # canonical == sorted(all, (-score, docid))[:k] has the same score sequence
    dc, sc, cc, _ = oracle.score_topk(dat, idx, c.n_docs, queries, 1000, tie_mode="canonical")
    for i, q in enumerate(g["queries"]):
        assert sc[i, :cc[i]].tolist() == [p[1] for p in q["top1000"]]
        assert cc[i] == min(1000, q["n_touched"])


# ------------------------------------------------------------------ reader + scoring
@pytest.mark.parametrize("name", ["kat", "zeros"])
def test_term_docs(golden, name):
    g = golden(name)
    idx = np.frombuffer(g["idx"], dtype=np.uint64)
    dat = np.frombuffer(g["dat"], dtype=np.uint8)
    for term, expect in g["term_docs"].items():
        if term not in g["vocab"]:
            assert expect == []
            continue
        t = g["vocab"].index(term)
        docs, vals = oracle.term_docs(dat, int(idx[2 * t]), int(idx[2 * t + 1]))
        assert [[int(a), int(b)] for a, b in zip(docs, vals)] == expect


@pytest.mark.parametrize("name", ["kat", "zeros"])
def test_score_raw_small_cases(golden, name):
    g = golden(name)
    idx = np.frombuffer(g["idx"], dtype=np.uint64)
    dat = np.frombuffer(g["dat"], dtype=np.uint8)
    tid = {t: i for i, t in enumerate(g["vocab"])}
    n_docs = len(g["lines"])
    for case in g["scores"]:
        k = case.get("top_k", 10)
        q = [[tid.get(t, -1) for t in case["terms"]]]
        d, s, cnt, _ = oracle.score_topk(dat, idx, n_docs, q, k, tie_mode="raw")
        assert as_pairs(d[0], s[0], cnt[0]) == case["result"]
        py = oracle.py_score(tid, g["dat"], g["idx"], case["terms"], k, canonical=False)
        assert [list(x) for x in py] == case["result"]


def test_score_small_collection(golden):
    g = golden("small")
    idx = np.frombuffer(g["idx"], dtype=np.uint64)
    dat = np.frombuffer(g["dat"], dtype=np.uint8)
    tid = {t: i for i, t in enumerate(g["vocab"])}
    queries = [[tid.get(t, -1) for t in q["terms"]] for q in g["queries"]]
    for key, k in (("all", 10 ** 6), ("top10", 10), ("top1", 1)):
        kk = min(k, g["n_docs"])
        d, s, cnt, _ = oracle.score_topk(dat, idx, g["n_docs"], queries, kk, tie_mode="raw")
        for i, q in enumerate(g["queries"]):
            assert as_pairs(d[i], s[i], cnt[i]) == q[key], (key, i)
    # canonical list == reference's full list re-sorted by (-score, docid), then cut
    for k in (1, 10, 200):
        d, s, cnt, _ = oracle.score_topk(dat, idx, g["n_docs"], queries, k, tie_mode="canonical")
        for i, q in enumerate(g["queries"]):
            canon = sorted(q["all"], key=lambda x: (-x[1], x[0]))[:k]
            assert as_pairs(d[i], s[i], cnt[i]) == canon
            py = oracle.py_score(tid, g["dat"], g["idx"], q["terms"], k, canonical=True)
            assert [list(x) for x in py] == canon


def test_csr_scorer_matches_file_scorer(golden):
    g = golden("small")
    t, v, o = parse_quantized(g["quantized_lines"], g["vocab"])
    toff, docs, imps = oracle.invert(t, v, o, len(g["vocab"]))
    dat, idx = oracle.serialize(toff, docs, imps)
    tid = {t: i for i, t in enumerate(g["vocab"])}
    queries = [[tid.get(t, -1) for t in q["terms"]] for q in g["queries"]]
    a = oracle.score_topk(dat, idx, g["n_docs"], queries, 50)
    b = oracle.score_topk_csr(toff, docs, imps, g["n_docs"], queries, 50)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_sparse_search_twin(golden):
    """SparseSearch.search (nano_beir_evaluator.py:103-137) == CSR scorer over dense doc ids."""
    g = golden("sparse")
    corpus_ids = list(g["corpus"].keys())
    term_of = {}
    lists = {}
    for dense, cid in enumerate(corpus_ids):
        for term, score in g["replay"][g["corpus"][cid]]:
            if score > 0:
                lists.setdefault(term_of.setdefault(term, len(term_of)), []).append((dense, int(score)))
    n_terms = len(term_of)
    toff = np.zeros(n_terms + 1, dtype=np.uint64)
    for t, l in lists.items():
        toff[t + 1] = len(l)
    toff = np.cumsum(toff).astype(np.uint64)
    docs = np.concatenate([np.array([d for d, _ in lists[t]], dtype=np.uint32) for t in range(n_terms)])
    imps = np.concatenate([np.array([s for _, s in lists[t]], dtype=np.uint8) for t in range(n_terms)])
    qids = list(g["queries"].keys())
    queries = [[term_of.get(t, -1) for t in dict.fromkeys(g["queries"][q].split())] for q in qids]
    for key, k in (("k10", 10), ("k1000", 1000)):
        d, s, cnt, _ = oracle.score_topk_csr(toff, docs, imps, len(corpus_ids), queries, k, tie_mode="raw")
        for i, q in enumerate(qids):
            got = [[corpus_ids[int(a)], float(b)] for a, b in zip(d[i, :cnt[i]], s[i, :cnt[i]])]
            assert got == g[key][q], (key, q)


def test_threshold_seed_is_a_lower_bound_of_the_kth_score():
    """The claim behind the GPU path's threshold seeds (csrc/build.cuh), checked on the CPU against the oracle:
    a query's k-th best score is at least the k-th highest impact of any single one of its terms — as long as no
    posting list names a document twice (the counter-example below is why such an index gets no seeds)."""
    from helpers import quantized_csr
    from improving_learned_index_b200 import synthetic as syn
    x = quantized_csr(3000, 120, 40, 5)
    toff = x["toff"].astype(np.int64)
    queries = syn.make_queries(60, vocab_size=120, seed=6)
    for k in (1, 10, 200, 2500):
        d, s, c, _ = oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], 3000, queries, k)
        for qi, q in enumerate(queries):
            bound = 0
            for t in set(q):
                imp = np.sort(x["vals"][toff[t]:toff[t + 1]])[::-1]
                if imp.size >= k:
                    bound = max(bound, int(imp[k - 1]))
            if bound:
                assert c[qi] == k and s[qi, k - 1] >= bound, (k, qi)
    # one term, document 5 listed twice with impact 200, every other document once with impact 1: two postings
    # reach 200 but only ONE document does; the 2nd best score is 1
    docs = np.array([5, 5] + [d for d in range(50) if d != 5], dtype=np.uint32)
    vals = np.array([200, 200] + [1] * 49, dtype=np.uint8)
    d, s, c, _ = oracle.score_topk_csr(np.array([0, 51], dtype=np.uint64), docs, vals, 50, [[0]], 2)
    assert s[0, :2].tolist() == [400, 1] and np.sort(vals)[::-1][1] == 200
