"""GPU tests of the API edges: argument validation, empty shards, file-based shards, pairwise queries,
large vocabularies, the quantize_max option of SparseSearch."""
import pickle

import numpy as np
import pytest

from improving_learned_index_b200 import _native, engine, synthetic as syn
from improving_learned_index_b200 import InvertedIndex, InvertedIndexCreator
from improving_learned_index_b200.evaluation import NanoBEIREvaluator, Ranker, SparseSearch
from improving_learned_index_b200.evaluation.nano_beir_evaluator import Dataset
from oracle import oracle
from helpers import assert_same_results, quantized_csr, write_index_dir

pytestmark = pytest.mark.gpu


def test_argument_validation():
    x = quantized_csr(500, 100, 20, 3)
    index = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=256)
    for bad_k in (0, 65537):
        with pytest.raises(_native.NativeError) as e:
            index.search([[1, 2]], bad_k)
        assert e.value.code == 2
    for bad_tile in (100, 255, 65536, 3000):
        with pytest.raises(_native.NativeError) as e:
            engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], tile_docs=bad_tile)
        assert e.value.code == 2
    with pytest.raises(ValueError):
        engine.DeviceIndex.from_csr(x["toff"], x["docs"][:-1], x["vals"])
    with pytest.raises(_native.NativeError):                       # doc range must be non-empty
        engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], doc_lo=10, doc_hi=10)
    with pytest.raises(TypeError):
        pickle.dumps(index)
    with pytest.raises(_native.NativeError) as e:                  # malformed .dat image
        engine.DeviceIndex.from_files(np.zeros(7, dtype=np.uint8), np.array([0, 10], dtype=np.uint64))
    assert e.value.code == 6
    with pytest.raises(_native.NativeError):                       # term id out of range in a collection
        engine.invert([0, 7], [1, 1], [0, 2], 3)


def test_empty_index_and_empty_shard():
    empty = engine.DeviceIndex.from_csr(np.zeros(5, dtype=np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint8))
    d, s, c = empty.search([[0, 1], []], 10)
    assert c.tolist() == [0, 0] and empty.info()["n_tiles"] == 0
    x = quantized_csr(1000, 100, 20, 4)
    shard = engine.DeviceIndex.from_csr(x["toff"], x["docs"], x["vals"], doc_lo=5000, doc_hi=6000)   # beyond every docid
    assert shard.info()["n_postings"] == 0
    assert shard.search([[1, 2, 3]], 5)[2].tolist() == [0]
    assert engine.find_max([]) == 0.0 and engine.quantize([], 3.0).size == 0


def test_file_shards_cover_the_index(golden, tmp_path):
    g = golden("small")
    path = write_index_dir(tmp_path / "small", g["vocab"], g["idx"], g["dat"])
    whole = InvertedIndex(path, tile_docs=256)
    parts = [InvertedIndex(path, doc_lo=lo, doc_hi=hi, tile_docs=256) for lo, hi in ((0, 70), (70, 150), (150, 200))]
    assert sum(p.device_index.info()["n_postings"] for p in parts) == whole.device_index.info()["n_postings"]
    for q in g["queries"]:
        merged = sorted((pair for p in parts for pair in p.score(q["terms"], top_k=200)), key=lambda x: (-x[1], x[0]))
        assert merged == whole.score(q["terms"], top_k=200)


def test_large_vocabulary_and_long_pairwise_queries(tmp_path):
    """A 200K-term vocabulary (pairwise 'a|b' terms blow the vocabulary up) and a pairwise Ranker run."""
    rng = np.random.default_rng(5)
    n_docs, V = 3000, 200_000
    terms = rng.integers(0, V, size=(n_docs, 12))
    lines = []
    for d in range(n_docs):
        uniq = dict.fromkeys(terms[d].tolist())
        lines.append(', '.join(f"w{t}: {1 + (t * 7 + d) % 200}" for t in uniq))
    a, b = int(terms[0][0]), int(terms[0][1])
    lines[0] += f", w{a}|w{b}: 77"
    src = tmp_path / "c"
    src.write_text(''.join(l + '\n' for l in lines))
    InvertedIndexCreator(src, tmp_path / "index").run()
    index = InvertedIndex(tmp_path / "index", tile_docs=512)
    assert len(index.vocab) > 30_000
    qfile = tmp_path / "q.tsv"
    qfile.write_text(f"1\tw{a} w{b}\n2\tw{int(terms[5][0])}\n")
    run = tmp_path / "run.tsv"
    Ranker(tmp_path / "index", qfile, run, pairwise=True, query_processor=lambda s: s.split(), top_k=5).run()
    rows = [l.split('\t') for l in run.read_text().split('\n')[:-1]]
    top = [r for r in rows if r[0] == "1"][0]
    vals = {t: 1 + (t * 7) % 200 for t in (a, b)}
    assert top[1] == "0" and int(top[3]) == (vals[a] + vals[b] + 77 if a != b else vals[a] + 77)
    # a 300-term query (32-bit accumulators, several rounds of 32 segments) against the oracle
    toff = np.fromfile(tmp_path / "index" / "inverted_index.idx", dtype=np.uint64)
    dat = np.fromfile(tmp_path / "index" / "inverted_index.dat", dtype=np.uint8)
    vocab = list(index.vocab)
    long_q = [vocab[i] for i in rng.integers(0, len(vocab), size=300)]
    ids = [index.vocab[t] for t in long_q]
    want = oracle.score_topk(dat, toff, n_docs, [ids], 50)
    got = index.score(long_q, top_k=50)
    assert got == list(zip(want[0][0, :want[2][0]].tolist(), want[1][0, :want[2][0]].tolist()))


def test_sparse_search_quantize_max_and_evaluator():
    class Model:
        def __init__(self, docs):
            self.docs = docs

        def get_impact_scores_batch(self, texts):
            return [self.docs[t] for t in texts]

        def process_query(self, query):
            return set(query.split())
    rng = np.random.default_rng(1)
    docs, corpus = {}, {}
    for d in range(400):
        text = f"doc {d}"
        corpus[f"d{d}"] = text
        docs[text] = [(f"t{t}", float(np.round(rng.uniform(0, 6.0), 3))) for t in rng.choice(50, size=8, replace=False)]
    queries = {f"q{i}": ' '.join(f"t{t}" for t in rng.choice(50, size=3, replace=False)) for i in range(20)}
    res = SparseSearch(Model(docs), 32, quantize_max=6.0).search(queries, corpus, k=10)
    # same thing by hand: reference quantize rule on every impact, then exact integer scoring
    for qid, text in queries.items():
        qt = set(text.split())
        scores = {}
        for cid, doc_text in corpus.items():
            s = sum(int(v * (255 / 6.0)) for t, v in docs[doc_text] if t in qt and int(v * (255 / 6.0)) > 0)
            if s:
                scores[cid] = float(s)
        order = list(corpus)
        want = sorted(scores.items(), key=lambda kv: (-kv[1], order.index(kv[0])))[:10]
        assert list(res[qid].items()) == want
    qrels = {q: {next(iter(r)): 1} for q, r in res.items() if r}
    ev = NanoBEIREvaluator(batch_size=32, datasets={"toy": Dataset(queries, corpus, qrels, "Toy")})

    class QModel(Model):          # integer impacts, as the default (strict) SparseSearch requires
        def get_impact_scores_batch(self, texts):
            return [[(t, float(int(v * (255 / 6.0)))) for t, v in self.docs[x]] for x in texts]
    out = ev.evaluate_all(QModel(docs))
    assert set(out) == {"toy", "avg"} and out["toy"][0]["NDCG@10"] == 1.0 and out["avg"][2]["Recall@10"] == 1.0
    # the reference's entry point with a model that emits FLOAT impacts (models/original.py:309): the evaluator
    # quantizes with the collection's own maximum (quantize.py:17-24,37) instead of refusing
    out_f = ev.evaluate_all(Model(docs))
    mx = max(v for lst in docs.values() for _, v in lst)
    by_hand = SparseSearch(Model(docs), 32, quantize_max=mx).search(queries, corpus, k=1000)
    auto = SparseSearch(Model(docs), 32, quantize_max="auto").search(queries, corpus, k=1000)
    assert auto == by_hand and set(out_f) == {"toy", "avg"} and 0.0 < out_f["toy"][0]["NDCG@10"] <= 1.0
    with pytest.raises(ValueError):
        SparseSearch(Model(docs), 32).search(queries, corpus, k=10)          # strict default: no silent rescaling
