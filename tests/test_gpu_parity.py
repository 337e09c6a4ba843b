"""GPU parity tests: every call goes through the C ABI (libdi_b200.so) and is compared bit for bit
with (a) the golden vectors recorded from the reference and (b) the CPU oracle on seeded inputs."""
import hashlib
import struct

import numpy as np
import pytest

from improving_learned_index_b200 import engine, synthetic as syn
from improving_learned_index_b200 import InvertedIndex, InvertedIndexCreator, quantize_file, find_max_value
from improving_learned_index_b200.evaluation import Metrics, Ranker, SparseSearch
from oracle import oracle
from helpers import assert_same_results, canonical, quantized_csr, write_index_dir

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ K1 quantize
def test_quantize_golden_cases(golden):
    g = golden("quantize")
    for case in g["cases"]:
        got = engine.quantize(case["values"], case["max"])
        assert got.tolist() == case["quantized"], case["max"]
        assert engine.find_max(case["values"]) == max(case["values"])


def test_quantize_self_max_sweep(golden):
    """max * (255 / max) lands on 254 for exactly the maxima the reference says (fp64, truncation)."""
    g = golden("quantize")
    vals = np.arange(1, 20001) / 1000
    got = np.array([engine.quantize([v], v)[0] for v in vals[::37]])
    want = np.array([oracle.quantize([v], v)[0] for v in vals[::37]])
    assert np.array_equal(got, want)
    sel = np.array(g["sweep_254"][:200])
    assert all(engine.quantize([m / 1000], m / 1000)[0] == 254 for m in sel[:50])


def test_quantize_random_matches_oracle():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.lognormal(0, 2, 200_000), rng.uniform(0, 1e-3, 1000), [0.0, 1e4, 5e-324]])
    x = np.round(x, 3)
    for mx in (None, 7.0, 0.013):
        assert np.array_equal(engine.quantize(x, mx).astype(np.int64),
                              np.clip(oracle.quantize(x, mx), -2**31, 2**31 - 1))
    assert engine.find_max(x) == oracle.find_max(x)
    assert engine.find_max([]) == 0.0 and engine.find_max([-1.0, -2.0]) == 0.0


def test_quantize_file_golden(golden, tmp_path):
    g = golden("quantize")["file"]
    src = tmp_path / "in"
    src.write_text(''.join(l + '\n' for l in g["lines"]))
    assert find_max_value(src) == 1.0
    for key, mx in (("auto", None), ("max2", 2.0), ("max05", 0.5)):
        quantize_file(src, tmp_path / key, mx)
        assert (tmp_path / key).read_text().split('\n')[:-1] == g[key]
    bad = tmp_path / "blank"
    bad.write_text("a: 1.0\n\nb: 2.0\n")
    with pytest.raises(ValueError):          # quantize.py:22,43 — blank line cannot be unpacked
        quantize_file(bad, tmp_path / "out")


# ------------------------------------------------------------------ K2 inversion, byte-identical files
@pytest.mark.parametrize("name", ["kat", "small", "zeros"])
def test_creator_bytes_identical(golden, tmp_path, name):
    g = golden(name)
    lines = g["quantized_lines"] if "quantized_lines" in g else g["lines"]
    src = tmp_path / "collection"
    src.write_text(''.join(l + '\n' for l in lines), encoding='utf-8')
    InvertedIndexCreator(src, tmp_path / "index").run()
    assert (tmp_path / "index" / "vocab.txt").read_text(encoding='utf-8').split('\n')[:-1] == g["vocab"]
    assert (tmp_path / "index" / "inverted_index.dat").read_bytes() == g["dat"]
    assert (tmp_path / "index" / "inverted_index.idx").read_bytes() == g["idx"]


def test_full_pipeline_medium_golden(golden, tmp_path):
    """raw text -> quantize_file -> InvertedIndexCreator -> InvertedIndex.score, all vs the reference's own run."""
    g = golden("medium")
    c = syn.make_collection(**{k: g["gen"][k] for k in ("n_docs", "vocab_size", "draws_per_doc", "seed")})
    raw = tmp_path / "collection.index"
    raw.write_text(''.join(l + '\n' for l in c.lines()), encoding='utf-8')
    quantize_file(raw, tmp_path / "q")
    sha = lambda b: hashlib.sha256(b).hexdigest()
    assert sha((tmp_path / "q").read_bytes()) == g["quantized_sha256"]
    InvertedIndexCreator(tmp_path / "q", tmp_path / "index").run()
    assert sha((tmp_path / "index" / "vocab.txt").read_bytes()) == g["vocab_sha256"]
    assert sha((tmp_path / "index" / "inverted_index.dat").read_bytes()) == g["dat_sha256"]
    assert sha((tmp_path / "index" / "inverted_index.idx").read_bytes()) == g["idx_sha256"]
    for tile_docs in (0, 256, 1024):
        index = InvertedIndex(tmp_path / "index", tile_docs=tile_docs, cand_slack=1000 if tile_docs else 0)
        got = index.score_batch([q["terms"] for q in g["queries"]], top_k=1000)
        for q, res in zip(g["queries"], got):
            ref = q["top1000"]
            # §8a(3): same score sequence; same doc set once boundary ties are removed
            assert [s for _, s in res] == [s for _, s in ref]
            kth = ref[-1][1] if len(ref) == 1000 else -1
            assert {d for d, s in res if s != kth} == {d for d, s in ref if s != kth}
            assert res == sorted(res, key=lambda x: (-x[1], x[0]))
            assert len(res) == min(1000, q["n_touched"])


def test_creator_rejects_out_of_range_impacts(tmp_path):
    src = tmp_path / "c"
    src.write_text("a: 3, b: 256\n")
    with pytest.raises(struct.error):
        InvertedIndexCreator(src, tmp_path / "i").run()


def test_invert_random_matches_oracle():
    for n_docs, V, draws, seed in ((1, 5, 3, 0), (700, 300, 30, 1), (20_000, 5000, 50, 2)):
        x = quantized_csr(n_docs, V, draws, seed)
        toff, docs, vals = engine.invert(x["terms"], x["imps"], x["offs"], V)
        assert np.array_equal(toff, x["toff"]) and np.array_equal(docs, x["docs"]) and np.array_equal(vals, x["vals"])
        dat, idx = engine.serialize(toff, docs, vals)
        dat_o, idx_o = oracle.serialize(x["toff"], x["docs"], x["vals"])
        assert np.array_equal(dat, dat_o) and np.array_equal(idx, idx_o)
    # empty collection
    toff, docs, vals = engine.invert([], [], [0, 0, 0], 4)
    assert toff.tolist() == [0] * 5 and docs.size == 0


def test_invert_is_stable_over_many_sort_blocks_and_flags_bad_term_ids():
    """K2's one-sweep sort: many 8192-key blocks per digit run (look-back chains), skewed digits (one hot term),
    a digit pass that is uniform (tiny vocabulary -> skipped on the device), term ids outside the vocabulary."""
    rng = np.random.default_rng(7)
    for n_docs, V, per_doc in ((60_000, 3, 3), (30_000, 700, 40), (9_000, 70_000, 25)):
        terms = np.concatenate([np.sort(rng.choice(V, size=min(per_doc, V), replace=False)) for _ in range(n_docs)]).astype(np.uint32)
        if V == 700:
            terms[rng.random(terms.size) < 0.5] = 5                      # half of all postings in one list (duplicates allowed)
        offs = np.arange(n_docs + 1, dtype=np.uint64) * min(per_doc, V)
        imps = rng.integers(0, 256, size=terms.size).astype(np.uint8)
        toff, docs, vals = engine.invert(terms, imps, offs, V)
        o_toff, o_docs, o_vals = oracle.invert(terms, imps, offs, V)
        assert np.array_equal(toff, o_toff) and np.array_equal(docs, o_docs) and np.array_equal(vals, o_vals), V
    with pytest.raises(RuntimeError):                                    # term id == n_terms: DI_ERR_RANGE
        engine.invert([0, 3], [1, 1], [0, 2], 3)


@pytest.mark.parametrize("n_docs,V,draws,tile_docs,dense_ratio", [(5000, 800, 100, 512, 0), (40_000, 2000, 60, 4096, 0),
                                                                    (40_000, 2000, 60, 256, 0xFFFFFFFF), (70_000, 3000, 40, 0, 0),
                                                                    (3000, 40, 30, 1024, 1)])
def test_index_from_docmajor_equals_index_from_csr(n_docs, V, draws, tile_docs, dense_ratio):
    """di_index_create_docmajor_dev (one segmented two-pass sort straight from the collection) must give the very
    index the term-major route gives: same statistics, same results as the oracle; shards by document range too."""
    torch = pytest.importorskip("torch")
    x = quantized_csr(n_docs, V, draws, 17)
    dev = torch.device("cuda:0")
    imps = x["imps"].copy()
    imps[::97] = 0                                                       # zero impacts are written by create.py and never read back
    o_toff, o_docs, o_vals = oracle.invert(x["terms"], imps, x["offs"], V)
    d_terms = torch.from_numpy(x["terms"].astype(np.int64)).to(dev).to(torch.int32)
    d_imps = torch.from_numpy(imps).to(dev)
    d_offs = torch.from_numpy(x["offs"].astype(np.int64)).to(dev)
    ref = engine.DeviceIndex.from_csr(o_toff, o_docs, o_vals, tile_docs=tile_docs, dense_ratio=dense_ratio)
    got = engine.DeviceIndex.from_docmajor_device(d_terms, d_imps, d_offs, n_docs, V, x["terms"].size,
                                                  tile_docs=tile_docs, dense_ratio=dense_ratio)
    a, b = ref.info(), got.info()
    for key in ("n_postings", "payload_bytes", "table_bytes", "n_dense_segments", "n_sparse_segments", "n_dense_postings", "tile_docs"):
        assert a[key] == b[key], key
    queries = syn.make_queries(300, vocab_size=V, seed=18)
    queries[0], queries[1] = [], [V + 1]
    for k in (10, 1000):
        want = oracle.score_topk_csr(o_toff, o_docs, o_vals, n_docs, queries, k)
        assert_same_results(got.search(queries, k), want, f"docmajor k={k}")
        assert_same_results(ref.search(queries, k), want, f"csr k={k}")
    flat = np.asarray([t for q in queries for t in q if 0 <= t < V], dtype=np.uint32)
    assert np.array_equal(got.term_df(flat), ref.term_df(flat))
    # a shard = a slice of the documents with its own doc_lo; docids stay global
    lo, hi = n_docs // 3, n_docs // 3 + n_docs // 2
    p0, p1 = int(x["offs"][lo]), int(x["offs"][hi])
    shard = engine.DeviceIndex.from_docmajor_device(d_terms[p0:p1], d_imps[p0:p1], (d_offs[lo:hi + 1] - p0).contiguous(),
                                                    hi - lo, V, p1 - p0, doc_lo=lo, tile_docs=tile_docs, dense_ratio=dense_ratio)
    twin = engine.DeviceIndex.from_csr(o_toff, o_docs, o_vals, doc_lo=lo, doc_hi=hi, tile_docs=tile_docs, dense_ratio=dense_ratio)
    assert_same_results(shard.search(queries, 100), twin.search(queries, 100), "shard")


# ------------------------------------------------------------------ reader semantics + scoring on golden indexes
@pytest.mark.parametrize("name", ["kat", "zeros"])
def test_index_golden_small(golden, tmp_path, name):
    g = golden(name)
    index = InvertedIndex(write_index_dir(tmp_path / name, g["vocab"], g["idx"], g["dat"]))
    for term, expect in g["term_docs"].items():
        assert [list(p) for p in index.term_docs(term)] == expect
    if "term_location" in g:
        for term, loc in g["term_location"].items():
            assert list(index.term_location(term)) == loc
    for case in g["scores"]:
        k = case.get("top_k", 10)
        full = index.score(case["terms"], top_k=10 ** 9)
        assert [list(p) for p in index.score(case["terms"], top_k=k)] == canonical(full, k)
        assert sorted(map(tuple, case["result"])) == sorted(canonical(full, 10 ** 9)[:len(case["result"])]) \
            or [s for _, s in case["result"]] == [s for _, s in canonical(full, k)]


@pytest.mark.parametrize("tile_docs,dense_ratio,cand_slack", [(0, 0, 0), (256, 0, 0), (256, 0xFFFFFFFF, 7), (256, 1, 1)])
def test_index_golden_small_collection(golden, tmp_path, tile_docs, dense_ratio, cand_slack):
    g = golden("small")
    index = InvertedIndex(write_index_dir(tmp_path / "small", g["vocab"], g["idx"], g["dat"]),
                          tile_docs=tile_docs, dense_ratio=dense_ratio, cand_slack=cand_slack)
    queries = [q["terms"] for q in g["queries"]]
    for k in (1, 3, 10, 200, 10 ** 9):
        got = index.score_batch(queries, top_k=k)
        for q, res in zip(g["queries"], got):
            assert [list(p) for p in res] == canonical(q["all"], k), (k, q["terms"])
    # single-query entry point, set input, generator input
    q = g["queries"][0]
    assert [list(p) for p in index.score(set(q["terms"]), top_k=5)] == canonical(q["all"], 5)
    assert [list(p) for p in index.score(iter(q["terms"]), top_k=5)] == canonical(q["all"], 5)


# ------------------------------------------------------------------ scoring vs oracle on seeded collections
CONFIGS = [
    # n_docs, V, draws, seed, tile_docs, dense_ratio, cand_slack, n_queries, ks
    (5000, 30522, 120, 10, 0, 0, 0, 50, (1000,)),                       # BASELINE config 1 shape
    (5000, 800, 100, 11, 1024, 0, 0, 80, (1, 10, 1000, 5000)),          # hot terms -> dense segments, multi-tile
    (5000, 800, 100, 11, 512, 0xFFFFFFFF, 64, 80, (7, 100)),            # all sparse, tiny candidate slack
    (5000, 800, 100, 11, 512, 1, 0, 80, (100,)),                        # everything dense
    (40_000, 2000, 60, 12, 4096, 0, 300, 120, (100, 1000)),
    (70_000, 3000, 40, 13, 0, 0, 0, 60, (10, 1000)),                    # default 16384-doc tiles, 5 tiles
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"n{c[0]}_V{c[1]}_t{c[4]}_d{c[5]}_s{c[6]}")
def test_search_matches_oracle(cfg):
    n_docs, V, draws, seed, tile_docs, dense_ratio, cand_slack, nq, ks = cfg
    x = quantized_csr(n_docs, V, draws, seed)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=tile_docs,
                                        dense_ratio=dense_ratio, cand_slack=cand_slack)
    info = index.info()
    assert info["n_postings"] == x["docs"].size
    queries = syn.make_queries(nq, vocab_size=V, seed=seed + 100)
    queries[0] = []                                   # empty query
    queries[1] = [V + 5, -1]                          # only out-of-vocabulary ids
    queries[2] = queries[3] + queries[3][:2]          # duplicated terms count again
    for k in ks:
        k = min(k, n_docs)
        got = index.search(queries, k)
        want = oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], n_docs, queries, k)
        assert_same_results(got, want, f"k={k}")
    flat = np.asarray([t for q in queries for t in q if 0 <= t < V], dtype=np.uint32)
    df = index.term_df(flat)
    assert np.array_equal(df, np.diff(x["toff"].astype(np.int64))[flat].astype(np.uint64))
    index.close()


@pytest.mark.parametrize("mode", ["seeded", "unseeded", "acc32", "per_tile"])
def test_long_tile_chains_with_one_lane_match_oracle(mode):
    """The configuration the bench times: MORE queries than resident CTAs, so every query is ONE chain of
    tiles (lanes == 1) handed from CTA to CTA through global memory (157 tiles here), with repeated cuts to k.
    Seeded and unseeded thresholds, 16- and 32-bit accumulators, and the one-launch-per-tile form."""
    from improving_learned_index_b200 import _native
    n_docs, V = 40_000, 2000
    x = quantized_csr(n_docs, V, 60, 71)
    flags = {"unseeded": _native.INDEX_NO_SEEDS, "per_tile": _native.INDEX_PER_TILE}.get(mode, 0)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=256, flags=flags)
    assert index.info()["n_tiles"] == 157
    queries = syn.make_queries(3000, vocab_size=V, seed=72)
    queries[0] = []
    queries[1] = [V + 9]
    queries[2] = queries[3] * 2
    hot = int(np.argmax(np.diff(x["toff"].astype(np.int64))))
    queries[4] = [hot]                                   # floods every tile until the threshold is exact
    if mode == "acc32":
        queries[5] = np.random.default_rng(3).integers(0, V, size=300).tolist()
    for k in ((1, 100, 1000) if mode == "seeded" else (100, 1000)):
        got = index.search(queries, k)
        t = index.timings()
        assert t["lanes"] == 1 and t["acc32"] == (1 if mode == "acc32" else 0), t
        assert t["score_launches"] == (157 if mode == "per_tile" else 1)
        want = oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], n_docs, queries, k)
        assert_same_results(got, want, f"{mode} k={k}")
    index.close()


def test_duplicate_postings_force_wide_accumulators():
    """A posting list may name a document several times (hand-made CSR, a model listing a term twice: the
    reference appends both). 200 query terms x 5 repeats x impact 255 = 255 000 overflows a u16 accumulator and
    exceeds the 255-per-term bound of the flood histogram: such an index must score with 32-bit accumulators."""
    n_docs, V, reps = 3000, 40, 5
    rng = np.random.default_rng(5)
    docs, vals, toff = [], [], [0]
    for t in range(V):
        d = np.sort(rng.choice(n_docs, size=1500, replace=False)).astype(np.uint32)
        v = rng.integers(200, 256, size=d.size).astype(np.uint8)
        if t < 8:                                        # document 7 repeated in the hot lists, impact 255
            d = np.concatenate([np.full(reps, 7, dtype=np.uint32), d[d != 7]])
            v = np.concatenate([np.full(reps, 255, dtype=np.uint8), v[:d.size - reps]])
        docs.append(d)
        vals.append(v)
        toff.append(toff[-1] + d.size)
    toff, docs, vals = np.asarray(toff, dtype=np.uint64), np.concatenate(docs), np.concatenate(vals)
    index = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=1024)
    queries = [list(range(8)) * 25, [0, 1, 2], list(range(V)) * 6, [3] * 257]
    for k in (1, 10, 2000):
        got = index.search(queries, k)
        assert index.timings()["acc32"] == 1
        assert_same_results(got, oracle.score_topk_csr(toff, docs, vals, n_docs, queries, k), f"k={k}")
    assert got[1][0, 0] >= 8 * 25 * reps * 255 and got[0][0, 0] == 7
    index.close()


@pytest.mark.parametrize("tile_docs", [256, 1024])
def test_tile_bounds_skip_tiles_and_keep_results_exact(tile_docs):
    """DI_INDEX_TILE_BOUNDS: a (query, tile) whose per-term impact maxima add up to less than the query's running
    threshold is skipped — a proof, so results must equal exhaustive scoring. Impacts fall with the docid here (a
    quality-ordered collection), so later tiles are provably empty of results and MUST be skipped; on the plain
    collection nothing can be skipped and nothing may change."""
    from improving_learned_index_b200 import _native
    n_docs, V = 40_000, 2000
    x = quantized_csr(n_docs, V, 60, 71)
    doc_of = x["docs"].astype(np.float64)
    skew = np.maximum(1, (x["vals"] * (1.0 - 0.97 * doc_of / n_docs)).astype(np.int64)).astype(np.uint8)
    queries = syn.make_queries(1500, vocab_size=V, seed=72)
    queries[0], queries[1] = [], [V + 9]
    queries[2] = queries[3] * 2
    queries[5] = np.random.default_rng(3).integers(0, V, size=40).tolist()      # more than one lookup round: never skipped
    for vals, expect_skips in ((skew, True), (x["vals"], False)):
        index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], vals, tile_docs=tile_docs, flags=_native.INDEX_TILE_BOUNDS)
        for k in (10, 1000):
            got = index.search(queries, k)
            t = index.timings()
            assert_same_results(got, oracle.score_topk_csr(x["toff"], x["docs"], vals, n_docs, queries, k), f"k={k}")
            n_pairs = len(queries) * index.info()["n_tiles"]
            if expect_skips:      # the tail of the collection cannot reach the threshold of a top-10 search
                assert t["tiles_skipped"] > (n_pairs // 20 if k == 10 else 0), (t["tiles_skipped"], n_pairs)
        one = index.search(queries[10:13], 10)                                  # tile lanes: every lane prunes from its seed on
        assert_same_results(one, oracle.score_topk_csr(x["toff"], x["docs"], vals, n_docs, queries[10:13], 10), "lanes")
        index.close()
    plain = engine.DeviceIndex.from_csr(x["toff"], x["docs"], skew, tile_docs=tile_docs)
    plain.search(queries, 10)
    assert plain.timings()["tiles_skipped"] == 0
    # the doc-major build keeps the same bounds
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    y = quantized_csr(20_000, 500, 40, 5)
    imps = np.maximum(1, (y["imps"] * (1.0 - 0.97 * np.repeat(np.arange(20_000), np.diff(y["offs"].astype(np.int64))) / 20_000)).astype(np.int64)).astype(np.uint8)
    o_toff, o_docs, o_vals = oracle.invert(y["terms"], imps, y["offs"], 500)
    dm = engine.DeviceIndex.from_docmajor_device(torch.from_numpy(y["terms"].astype(np.int64)).to(dev).to(torch.int32),
                                                 torch.from_numpy(imps).to(dev), torch.from_numpy(y["offs"].astype(np.int64)).to(dev),
                                                 20_000, 500, y["terms"].size, tile_docs=tile_docs, flags=_native.INDEX_TILE_BOUNDS)
    qs = syn.make_queries(1200, vocab_size=500, seed=9)
    assert_same_results(dm.search(qs, 100), oracle.score_topk_csr(o_toff, o_docs, o_vals, 20_000, qs, 100), "docmajor bounds")
    assert dm.timings()["tiles_skipped"] > 0


def test_long_queries_use_32bit_accumulators():
    """> 257 term occurrences can overflow a u16 accumulator; > 32 terms need several rounds."""
    x = quantized_csr(3000, 400, 120, 21)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=1024)
    rng = np.random.default_rng(0)
    queries = [rng.integers(0, 400, size=n).tolist() for n in (33, 64, 100, 257, 258, 300, 700)]
    queries.append([int(np.argmax(np.diff(x["toff"].astype(np.int64))))] * 300)      # one hot term 300 times
    for k in (10, 3000):
        assert_same_results(index.search(queries, k),
                            oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], 3000, queries, k), f"k={k}")
    short = [q[:40] for q in queries]
    assert_same_results(index.search(short, 50),
                        oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], 3000, short, 50), "short")


def test_massive_ties_are_docid_ordered():
    """One term, every document the same impact: top-k must be the k lowest docids, across tiles."""
    n = 3000
    toff = np.array([0, n, n + 3], dtype=np.uint64)
    docs = np.concatenate([np.arange(n, dtype=np.uint32)[::-1], np.array([5, 2999, 7], dtype=np.uint32)])
    vals = np.concatenate([np.full(n, 9, dtype=np.uint8), np.array([1, 1, 1], dtype=np.uint8)])
    for dense_ratio in (0, 0xFFFFFFFF):
        index = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=256, dense_ratio=dense_ratio, cand_slack=16)
        d, s, c = index.search([[0], [0, 1], [1]], 10)
        assert c.tolist() == [10, 10, 3]
        assert d[0, :10].tolist() == list(range(10)) and set(s[0, :10].tolist()) == {9}
        assert d[1, :3].tolist() == [5, 7, 2999] and s[1, :3].tolist() == [10, 10, 10]
        assert d[2, :3].tolist() == [5, 7, 2999]


def test_duplicate_postings_and_hidden_zeros():
    """Hand-made CSR: a (term, doc) pair stored twice counts twice; postings after the first zero
    impact of a list are invisible (inverted_index.py:50-51) even if non-zero ones follow."""
    toff = np.array([0, 4, 8], dtype=np.uint64)
    docs = np.array([3, 3, 1, 2, 0, 1, 2, 3], dtype=np.uint32)
    vals = np.array([5, 6, 7, 8, 4, 0, 9, 9], dtype=np.uint8)
    for dense_ratio in (1, 0xFFFFFFFF):
        index = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=256, dense_ratio=dense_ratio)
        assert index.info()["n_postings"] == 5
        d, s, c = index.search([[0], [1], [0, 1]], 4)
        assert list(zip(d[0, :c[0]].tolist(), s[0, :c[0]].tolist())) == [(3, 11), (2, 8), (1, 7)]
        assert list(zip(d[1, :c[1]].tolist(), s[1, :c[1]].tolist())) == [(0, 4)]
        want = oracle.score_topk_csr(toff, docs, vals, 4, [[0, 1]], 4)
        assert d[2, :c[2]].tolist() == want[0][0, :want[2][0]].tolist()


@pytest.mark.parametrize("tile_docs,cand_slack", [(0, 0), (1024, 0), (512, 40)])
def test_threshold_seeds_keep_results_exact(tile_docs, cand_slack):
    """Frequent terms (df >= 4096) get a per-term impact table and every query starts from a proven lower
    bound on its k-th best score (build.cuh, threshold seeds). Tiny vocabulary = every term is frequent and the
    first tiles flood the candidate lists; results must still equal exhaustive scoring for every k."""
    n_docs, V = 30_000, 60
    x = quantized_csr(n_docs, V, 30, 61)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=tile_docs, cand_slack=cand_slack)
    queries = syn.make_queries(70, vocab_size=V, seed=62)
    queries[0] = [int(np.argmax(np.diff(x["toff"].astype(np.int64))))]          # the hottest term alone
    queries[1] = queries[0] * 3                                                  # ... three times
    queries[2] = []
    for k in (1, 50, 1000, 4500, n_docs):
        assert_same_results(index.search(queries, k),
                            oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], n_docs, queries, k), f"k={k}")
    one = [queries[5]]                                                           # a single query runs in tile lanes
    assert_same_results(index.search(one, 300), oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], n_docs, one, 300), "lanes")
    index.close()


def test_duplicate_document_in_a_frequent_list_gets_no_seed():
    """k postings with impact >= v prove k DOCUMENTS only if no document is listed twice. Term 0 lists
    document 5 twice with impact 200 and 5000 other documents with impact 1: the 2nd best score is 1, not 200."""
    n = 5001
    others = np.array([d for d in range(n) if d != 5], dtype=np.uint32)
    toff = np.array([0, n + 1], dtype=np.uint64)
    docs = np.concatenate([np.array([5, 5], dtype=np.uint32), others])
    vals = np.concatenate([np.array([200, 200], dtype=np.uint8), np.ones(n - 1, dtype=np.uint8)])
    index = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=1024)
    for k in (1, 2, 3, 100):
        got = index.search([[0]], k)
        assert_same_results(got, oracle.score_topk_csr(toff, docs, vals, n, [[0]], k), f"k={k}")
    d, s, c = index.search([[0]], 3)
    assert d[0, :3].tolist() == [5, 0, 1] and s[0, :3].tolist() == [400, 1, 1]
    # without the duplicate the bound is used and must be just as exact
    toff2 = np.array([0, n], dtype=np.uint64)
    index2 = engine.DeviceIndex.from_csr(toff2, docs[1:], vals[1:], tile_docs=1024)
    d, s, c = index2.search([[0]], 3)
    assert d[0, :3].tolist() == [5, 0, 1] and s[0, :3].tolist() == [200, 1, 1]


def test_shards_and_merge_equal_single_index():
    """K5: docid-range shards searched separately + merged == one index (single GPU, 3 shards)."""
    torch = pytest.importorskip("torch")
    x = quantized_csr(9000, 1500, 60, 31)
    full = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=1024)
    queries = syn.make_queries(40, vocab_size=1500, seed=5)
    k = 100
    want = full.search(queries, k)
    bounds = [(0, 2500), (2500, 7000), (7000, 9000)]
    flat, offs = engine.flatten_queries(queries)
    dev = torch.device("cuda:0")
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)  # bit pattern of u32 ids
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    keys = torch.zeros((len(bounds), len(queries), k), dtype=torch.int64, device=dev)
    counts = torch.zeros((len(bounds), len(queries)), dtype=torch.int32, device=dev)
    shards = []
    for s, (lo, hi) in enumerate(bounds):
        shard = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], doc_lo=lo, doc_hi=hi, tile_docs=1024)
        shards.append(shard)
        assert shard.info()["doc_lo"] == lo
        shard.search_device(d_flat, d_offs, len(queries), max(len(q) for q in queries), k, keys[s], counts[s],
                            torch.cuda.current_stream().cuda_stream)
    out_keys = torch.zeros((len(queries), k), dtype=torch.int64, device=dev)
    out_counts = torch.zeros(len(queries), dtype=torch.int32, device=dev)
    engine.merge_topk_device(keys, counts, len(bounds), len(queries), k, out_keys, out_counts,
                             torch.cuda.current_stream().cuda_stream)
    docids = torch.zeros((len(queries), k), dtype=torch.int32, device=dev)
    scores = torch.zeros((len(queries), k), dtype=torch.int32, device=dev)
    engine.unpack_keys_device(out_keys, out_keys.numel(), docids, scores, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = (docids.cpu().numpy().view(np.uint32), scores.cpu().numpy(), out_counts.cpu().numpy().view(np.uint32))
    assert_same_results(got, want, "merged")
    assert sum(s.info()["n_postings"] for s in shards) == full.info()["n_postings"]


@pytest.mark.parametrize("k,p", [(1000, 221), (1000, 1), (300, 120), (64, 64), (1000, 999), (100, 7)])
def test_sorted_prefix_rows_hold_the_same_top_k(k, p):
    """di_index_set_sorted_prefix(p): a device row is [its p best keys, sorted | the rest of the top-k in any order] — the
    same SET as the fully sorted row and the same first p columns; the host API stays fully sorted."""
    torch = pytest.importorskip("torch")
    x = quantized_csr(40_000, 2000, 60, 71)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=1024)
    queries = syn.make_queries(1500, vocab_size=2000, seed=8)         # more than resident CTAs: one lane, the sharded case
    queries[0], queries[1] = [], [int(np.argmax(np.diff(x["toff"].astype(np.int64))))]
    flat, offs = engine.flatten_queries(queries)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    n, max_len = len(queries), max(len(q) for q in queries)

    def rows():
        keys = torch.zeros((n, k), dtype=torch.int64, device=dev)
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        index.search_device(d_flat, d_offs, n, max_len, k, keys, counts, st)
        torch.cuda.synchronize()
        return keys.cpu().numpy().view(np.uint64), counts.cpu().numpy()
    full, c_full = rows()
    index.set_sorted_prefix(p)
    part, c_part = rows()
    host = index.search(queries, k)                                   # host rows ignore the setting
    index.set_sorted_prefix(0)
    assert np.array_equal(c_full, c_part)
    for i in range(n):
        c = int(c_full[i])
        assert c < 2 or np.all(full[i, 1:c] < full[i, :c - 1])
        assert np.array_equal(part[i, :min(p, c)], full[i, :min(p, c)]), i
        assert np.array_equal(np.sort(part[i, :c]), np.sort(full[i, :c])), i
        assert np.array_equal((~(full[i, :c] & np.uint64(0xFFFFFFFF)).astype(np.uint32)), host[0][i, :c])
    index.close()


def test_global_seed_tables_make_shards_emit_less_and_merge_exactly():
    """di_index_export_seed_hist_dev / import: with the impact histograms of all shards added up, a shard's seeds bound the
    GLOBAL k-th score: its rows shrink to what can be in the global top-k, and the merged result is unchanged."""
    torch = pytest.importorskip("torch")
    n_docs, V, k = 120_000, 300, 100
    x = quantized_csr(n_docs, V, 40, 91)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    queries = syn.make_queries(1200, vocab_size=V, seed=92)
    flat, offs = engine.flatten_queries(queries)
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    Q, max_len = len(queries), max(len(q) for q in queries)
    want = oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], n_docs, queries, k)
    bounds = [(0, 30_000), (30_000, 60_000), (60_000, 90_000), (90_000, n_docs)]
    shards = [engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], doc_lo=lo, doc_hi=hi, tile_docs=4096) for lo, hi in bounds]

    def run():
        keys = torch.zeros((len(shards), Q, k), dtype=torch.int64, device=dev)
        counts = torch.zeros((len(shards), Q), dtype=torch.int32, device=dev)
        for s, shard in enumerate(shards):
            shard.search_device(d_flat, d_offs, Q, max_len, k, keys[s], counts[s], st)
        out_keys = torch.zeros((Q, k), dtype=torch.int64, device=dev)
        out_counts = torch.zeros(Q, dtype=torch.int32, device=dev)
        engine.merge_topk_device(keys, counts, len(shards), Q, k, out_keys, out_counts, st)
        torch.cuda.synchronize()
        kk = out_keys.cpu().numpy().view(np.uint64)
        return ((~(kk & np.uint64(0xFFFFFFFF)).astype(np.uint32)), (kk >> np.uint64(32)).astype(np.int32),
                out_counts.cpu().numpy().view(np.uint32)), int(counts.sum())
    got, emitted_local = run()
    assert_same_results(got, want, "local seeds")
    total = torch.zeros((V, 256), dtype=torch.int32, device=dev)
    for shard in shards:
        h = torch.zeros_like(total)
        shard.export_seed_hist(h, st)
        total += h
    torch.cuda.synchronize()
    assert int(total.sum()) == x["docs"].size and int(total[:, 0].sum()) == 0
    for shard in shards:
        shard.import_seed_hist(total, st)
    got, emitted_global = run()
    assert_same_results(got, want, "global seeds")
    assert emitted_global < emitted_local, (emitted_global, emitted_local)


def test_initial_thresholds_cut_the_result_exactly():
    """di_search_dev with caller-proven lower bounds returns exactly the keys at or above them."""
    torch = pytest.importorskip("torch")
    x = quantized_csr(9000, 1200, 60, 51)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=1024)
    queries = syn.make_queries(40, vocab_size=1200, seed=8)
    flat, offs = engine.flatten_queries(queries)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    k, max_len = 200, max(len(q) for q in queries)
    keys = torch.zeros((len(queries), k), dtype=torch.int64, device=dev)
    counts = torch.zeros(len(queries), dtype=torch.int32, device=dev)
    index.search_device(d_flat, d_offs, len(queries), max_len, k, keys, counts, st)
    torch.cuda.synchronize()
    cut = torch.clamp(counts.to(torch.int64) // 3, min=0)                      # keep about a third of every row
    theta = torch.where(counts > 0, keys[torch.arange(len(queries), device=dev), cut], torch.zeros_like(cut))
    keys2 = torch.zeros_like(keys)
    counts2 = torch.zeros_like(counts)
    index.search_device(d_flat, d_offs, len(queries), max_len, k, keys2, counts2, st, d_theta_init=theta.contiguous())
    torch.cuda.synchronize()
    for i in range(len(queries)):
        n = int(cut[i]) + 1 if int(counts[i]) else 0
        assert int(counts2[i]) == n, i
        assert torch.equal(keys2[i, :n], keys[i, :n]), i


def test_short_rows_merge_proves_or_flags():
    """K5 with k_in < k: unflagged queries must equal the single-index result; a query whose top-k sits in
    one shard must be flagged (and is exact again once re-run with full rows)."""
    torch = pytest.importorskip("torch")
    x = quantized_csr(6000, 900, 60, 41)
    # term 900 lives only in documents 0..599 (first shard)
    toff = np.concatenate([x["toff"], [x["toff"][-1] + 600]]).astype(np.uint64)
    docs = np.concatenate([x["docs"], np.arange(600, dtype=np.uint32)])
    vals = np.concatenate([x["vals"], (255 - np.arange(600) % 100).astype(np.uint8)])
    full = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=1024)
    queries = syn.make_queries(30, vocab_size=900, seed=6)
    queries[7] = [900]
    k, k_in = 300, 120
    want = full.search(queries, k)
    bounds = [(0, 2000), (2000, 4000), (4000, 6000)]
    flat, offs = engine.flatten_queries(queries)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    shards = [engine.DeviceIndex.from_csr(toff, docs, vals, doc_lo=lo, doc_hi=hi, tile_docs=1024) for lo, hi in bounds]

    def run(rows):
        keys = torch.zeros((3, len(queries), rows), dtype=torch.int64, device=dev)
        counts = torch.zeros((3, len(queries)), dtype=torch.int32, device=dev)
        for s, shard in enumerate(shards):
            shard.search_device(d_flat, d_offs, len(queries), max(len(q) for q in queries), rows, keys[s], counts[s], st)
        out_keys = torch.zeros((len(queries), k), dtype=torch.int64, device=dev)
        out_counts = torch.zeros(len(queries), dtype=torch.int32, device=dev)
        flags = torch.full((len(queries),), 7, dtype=torch.int32, device=dev)
        engine.merge_topk_device(keys, counts, 3, len(queries), k, out_keys, out_counts, st, k_in=rows, d_incomplete=flags)
        torch.cuda.synchronize()
        keys_np = out_keys.cpu().numpy().view(np.uint64)
        return ((~(keys_np & np.uint64(0xFFFFFFFF)).astype(np.uint32)), (keys_np >> np.uint64(32)).astype(np.int32),
                out_counts.cpu().numpy().view(np.uint32), flags.cpu().numpy())

    d, s, c, flags = run(k_in)
    assert flags[7] == 1 and set(flags.tolist()) <= {0, 1}
    for i in np.flatnonzero(flags == 0):
        n = int(want[2][i])
        assert c[i] == n and np.array_equal(d[i, :n], want[0][i, :n]) and np.array_equal(s[i, :n], want[1][i, :n]), i
    d, s, c, flags = run(k)
    assert not flags.any()
    assert_same_results((d, s, c), want, "full rows")


# ------------------------------------------------------------------ in-memory twin, ranker, metrics
def test_sparse_search_golden(golden):
    g = golden("sparse")

    class Replay:
        def get_impact_scores_batch(self, texts):
            return [[(t, np.float32(v)) for t, v in g["replay"][x]] for x in texts]

        def process_query(self, query):
            return set(query.split())
    corpus_pos = {cid: i for i, cid in enumerate(g["corpus"])}
    for key, k in (("k10", 10), ("k1000", 1000)):
        searcher = SparseSearch(Replay(), batch_size=16)
        res = searcher.search(g["queries"], g["corpus"], k)
        assert list(res.keys()) == list(g["queries"].keys())
        full = SparseSearch(Replay(), batch_size=50).search(g["queries"], g["corpus"], 10 ** 6)
        for qid, ref in g[key].items():
            got = list(res[qid].items())
            assert [s for _, s in got] == [s for _, s in ref], qid
            canon = sorted(full[qid].items(), key=lambda x: (-x[1], corpus_pos[x[0]]))[:k]
            assert got == canon, qid
            kth = ref[-1][1] if len(ref) == k else -1.0
            assert {d for d, s in got if s != kth} == {d for d, s in ref if s != kth}
            assert all(isinstance(s, float) for _, s in got)
        assert set(searcher.inverted_index.keys()) == {t for lst in g["replay"].values() for t, v in lst if v > 0}
    with pytest.raises(ValueError):
        class Frac(Replay):
            def get_impact_scores_batch(self, texts):
                return [[("a", 0.5)] for _ in texts]
        SparseSearch(Frac(), 4).search({"q": "a"}, {"d": "x"}, 1)


def test_ranker_and_metrics_end_to_end(golden, tmp_path):
    g = golden("small")
    index_dir = write_index_dir(tmp_path / "index", g["vocab"], g["idx"], g["dat"])
    qfile = tmp_path / "queries.tsv"
    rows = [(f"{100 + i}", ' '.join(q["terms"])) for i, q in enumerate(g["queries"]) if q["terms"]]
    qfile.write_text(''.join(f"{qid}\t{text}\n" for qid, text in rows))
    run = tmp_path / "run.tsv"
    Ranker(index_dir, qfile, run, num_workers=3, query_processor=lambda s: s.split(), top_k=10, batch_size=7).run()
    lines = run.read_text().split('\n')[:-1]
    by_q = {}
    for line in lines:
        qid, pid, rank, score = line.split('\t')
        by_q.setdefault(qid, []).append([int(pid), int(score)])
        assert int(rank) == len(by_q[qid])
    assert list(by_q) == [qid for qid, _ in rows if qid in by_q]            # batches reach the file in query order
    for (qid, _), q in zip(rows, [q for q in g["queries"] if q["terms"]]):
        terms = list(dict.fromkeys(q["terms"]))          # the ranker de-duplicates (set), like process_query
        want = oracle.py_score({t: i for i, t in enumerate(g["vocab"])}, g["dat"], g["idx"], terms, 10, canonical=True)
        assert by_q.get(qid, []) == [list(p) for p in want]
    # the same run with qrels: the report computed from the device-resident keys (di_eval_ranks_dev) must equal what
    # Metrics.evaluate reads back out of the run file, and the file must be the same bytes
    rng = np.random.default_rng(11)
    qrels = tmp_path / "qrels.tsv"
    with open(qrels, "w") as f:
        for qid, hits in by_q.items():
            picks = {hits[int(rng.integers(0, len(hits)))][0], hits[-1][0], 10 ** 6 + int(qid)}   # two retrieved, one never retrieved
            if int(qid) % 4 == 0:
                picks = {10 ** 6 + int(qid)}                                   # a query with no retrieved relevant passage
            for pid in sorted(picks):
                f.write(f"{qid}\t0\t{pid}\t1\n")
        f.write("99999\t0\t5\t1\n")                                           # in the qrels, not in the query file? -> must be in it
    qfile.write_text(qfile.read_text() + "99999\tzzz-not-in-vocab\n")
    run2 = tmp_path / "run2.tsv"
    depths = dict(mrr_depths=[1, 10], recall_depths=[1, 3, 5, 10])
    report = Ranker(index_dir, qfile, run2, qrels_path=qrels, query_processor=lambda s: s.split(), top_k=10, batch_size=5).run(**depths)
    want_report = Metrics(run2, qrels, **depths).evaluate()
    assert report == want_report and any(v > 0 for v in report.values())
    kept = [l for l in lines if l.split('\t')[0] in by_q]
    assert sorted(run2.read_text().split('\n')[:-1]) == sorted(kept)


def test_metrics_golden(golden, tmp_path):
    g = golden("metrics")
    (tmp_path / "run.tsv").write_text('\n'.join(g["run"]) + '\n')
    (tmp_path / "qrels.tsv").write_text('\n'.join(g["qrels"]) + '\n')
    m = Metrics(tmp_path / "run.tsv", tmp_path / "qrels.tsv", mrr_depths=[10, 100], recall_depths=[3, 10, 20, 50])
    rep = m.evaluate()
    assert {k.split('@')[1]: v for k, v in rep.items() if k.startswith('MRR')} == g["mrr"]
    assert {k.split('@')[1]: v for k, v in rep.items() if k.startswith('Recall')} == g["recall"]
