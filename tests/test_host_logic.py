"""CPU-only tests: ABI surface, host-side parsers / metrics, loud failure without a GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from improving_learned_index_b200 import _native, engine
from improving_learned_index_b200.evaluation.metrics import Metrics
from improving_learned_index_b200.evaluation.trec_metrics import EvaluateRetrieval
from improving_learned_index_b200.indexing.deep_impact_collection import (DeepImpactCollection,
                                                                         DeepPairwiseImpactCollection, parse_line)
from improving_learned_index_b200.utils.datasets import Queries, QueryParser, QueryRelevanceDataset, RunFile

REPO = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    """libdi_b200.so loads without a GPU and exports exactly what include/di_b200.h declares."""
    header = (REPO / "include" / "di_b200.h").read_text()
    declared = set(re.findall(r"\b(di_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    _native.build()
    lib = ctypes.CDLL(str(_native.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert _native.lib().di_version() >= 100


def test_no_silent_cpu_fallback():
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeError) as e:
        engine.quantize([1.0, 2.0])
    assert e.value.code == 5
    with pytest.raises(_native.NativeError):
        engine.DeviceIndex.from_csr([0, 1], [0], [1])


def test_product_code_never_touches_the_oracle():
    pkg = REPO / "improving-learned-index_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu*")):
        text = path.read_text()
        assert "oracle" not in text.lower() or path.name == "__init__.py", path


def test_flatten_queries():
    flat, offs = engine.flatten_queries([[1, 2], [], [-1, None, 7]])
    assert flat.tolist() == [1, 2, 0xFFFFFFFF, 0xFFFFFFFF, 7] and offs.tolist() == [0, 2, 2, 5]


def test_collection_parser(tmp_path):
    p = tmp_path / "c"
    p.write_text("a: 1.5, b: 2, a: 3\n\n  \nx|y: 4, x: 1, y: 2\n")
    c = DeepImpactCollection(p)
    assert len(c) == 4 and c[0] == {"a": 3.0, "b": 2.0} and c[1] == {} and c[2] == {}
    assert list(c[0]) == ["a", "b"]
    assert [pid for pid, _ in c] == [0, 1, 2, 3]
    assert c.score(0, {"a", "zz"}) == 3.0
    assert DeepPairwiseImpactCollection(p).score(3, {"x", "y"}) == 7.0
    with pytest.raises(ValueError):
        parse_line("a 1.5")


def test_query_and_run_files(tmp_path):
    q = tmp_path / "q.tsv"
    q.write_text("7\thello world\n8\tfoo\n")
    qs = Queries(q)
    assert len(qs) == 2 and qs[7] == "hello world" and list(qs.keys()) == ["7", "8"]
    assert QueryParser.parse('{"_id": "a1", "text": "t"}', 'beir') == ("a1", "t")
    (tmp_path / "qrels").write_text("7\t0\t11\t1\n7\t0\t12\t1\n8\t0\t5\t1\n")
    rel = QueryRelevanceDataset(tmp_path / "qrels")
    assert rel[7] == {"11", "12"} and len(rel) == 2
    (tmp_path / "bad").write_text("7\t1\t11\t1\n")
    with pytest.raises(AssertionError):
        QueryRelevanceDataset(tmp_path / "bad")
    (tmp_path / "empty").write_text("")
    with pytest.raises(ZeroDivisionError):           # datasets.py:161 averages over zero queries
        QueryRelevanceDataset(tmp_path / "empty")
    assert rel.average_positive_per_query == 1.5
    run = RunFile(tmp_path / "run")
    run.writelines("7", [(11, 300), (3, 200)])
    run.write("8", 5, 1, 10)
    run.writelines("7", [(12, 1)])               # append semantics, like the reference
    assert list(run.read()) == [("7", "11", 1, 300.0), ("7", "3", 2, 200.0), ("8", "5", 1, 10.0), ("7", "12", 1, 1.0)]


def test_metrics_against_reference_numbers(golden, tmp_path):
    g = golden("metrics")
    (tmp_path / "run.tsv").write_text('\n'.join(g["run"]) + '\n')
    (tmp_path / "qrels.tsv").write_text('\n'.join(g["qrels"]) + '\n')
    m = Metrics(tmp_path / "run.tsv", tmp_path / "qrels.tsv", mrr_depths=[10, 100], recall_depths=[3, 10, 20, 50])
    rep = m.evaluate()
    assert {k.split('@')[1]: v for k, v in rep.items() if k.startswith('MRR')} == g["mrr"]
    assert {k.split('@')[1]: v for k, v in rep.items() if k.startswith('Recall')} == g["recall"]


def test_trec_metrics_hand_computed():
    qrels = {"q1": {"d1": 1, "d3": 1}, "q2": {"d9": 1}}
    results = {"q1": {"d1": 3.0, "d2": 2.0, "d3": 1.0, "q1": 99.0}, "q2": {"d8": 1.0}}
    ndcg, _map, recall, prec = EvaluateRetrieval().evaluate(qrels, results, [1, 3])
    import math
    dcg3 = 1 + 1 / math.log2(4)
    idcg3 = 1 + 1 / math.log2(3)
    assert ndcg["NDCG@1"] == round((1.0 + 0.0) / 2, 5)
    assert ndcg["NDCG@3"] == round((dcg3 / idcg3 + 0.0) / 2, 5)
    assert _map["MAP@3"] == round(((1 + 2 / 3) / 2 + 0.0) / 2, 5)
    assert recall["Recall@1"] == 0.25 and recall["Recall@3"] == 0.5
    assert prec["P@1"] == 0.5 and prec["P@3"] == round((2 / 3) / 2, 5)


def test_trec_metrics_against_scikit_learn():
    """An independent, published implementation as the external check of trec_metrics.py (beir / pytrec_eval are not
    installable here): scikit-learn's ndcg_score (linear gain, log2 discount, ideal ranking from the true gains) and
    average_precision_score (sum of precision at the relevant ranks / number of relevant documents) on random graded
    qrels. sklearn sees the whole corpus (documents the run did not retrieve rank below every retrieved one), which is
    how trec_eval counts relevant-but-unretrieved documents in the ideal DCG and in the AP / recall denominators."""
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(5)
    n_docs, n_queries, depth = 400, 40, 150
    docs = [f"d{i:04d}" for i in range(n_docs)]
    qrels, results, truth, scores = {}, {}, {}, {}
    for qi in range(n_queries):
        q = f"q{qi}"
        gains = np.zeros(n_docs, dtype=np.int64)
        rel = rng.choice(n_docs, size=int(rng.integers(1, 30)), replace=False)
        gains[rel] = rng.integers(1, 4, size=rel.size)          # graded relevance 1..3
        judged_zero = rng.choice(n_docs, size=10, replace=False)  # explicit 0 judgements must not count
        sc = rng.permutation(n_docs).astype(np.float64) + 1.0    # distinct scores: no tie rule involved
        retrieved = np.argsort(-sc)[:depth]
        qrels[q] = {docs[d]: int(gains[d]) for d in rel}
        qrels[q].update({docs[d]: 0 for d in judged_zero if gains[d] == 0})
        results[q] = {docs[d]: float(sc[d]) for d in retrieved}
        full = np.full(n_docs, 0.0)
        full[retrieved] = sc[retrieved]                           # unretrieved: below every retrieved score
        full[np.setdiff1d(np.arange(n_docs), retrieved)] = -1.0 - np.arange(n_docs - depth)
        truth[q], scores[q] = gains, full
    ks = [1, 5, 10, 100]
    ndcg, _map, recall, prec = EvaluateRetrieval().evaluate(qrels, results, ks + [depth])
    for k in ks:
        want = np.mean([sk.ndcg_score(truth[q][None, :], scores[q][None, :], k=k) for q in qrels])
        assert ndcg[f"NDCG@{k}"] == round(float(want), 5)
        order = {q: np.argsort(-scores[q])[:k] for q in qrels}
        assert recall[f"Recall@{k}"] == round(float(np.mean([(truth[q][order[q]] > 0).sum() / (truth[q] > 0).sum() for q in qrels])), 5)
        assert prec[f"P@{k}"] == round(float(np.mean([(truth[q][order[q]] > 0).sum() / k for q in qrels])), 5)
    # MAP: sklearn has no cut-off; map_cut.k is the same sum restricted to ranks <= k, so compare on a run that holds every
    # relevant document above the cut (retrieve everything, cut at the corpus size)
    results_all = {q: {docs[d]: float(scores[q][d]) for d in range(n_docs)} for q in qrels}
    _, map_all, _, _ = EvaluateRetrieval().evaluate(qrels, results_all, [n_docs])
    want = np.mean([sk.average_precision_score(truth[q] > 0, scores[q]) for q in qrels])
    assert map_all[f"MAP@{n_docs}"] == round(float(want), 5)
    # and the cut version against a direct restatement of trec_eval's map_cut on the same ranking
    for k in ks:
        vals = []
        for q in qrels:
            order = np.argsort(-scores[q])[:k]
            hit = truth[q][order] > 0
            vals.append(float((np.cumsum(hit)[hit] / (np.flatnonzero(hit) + 1)).sum() / (truth[q] > 0).sum()))
        assert _map[f"MAP@{k}"] == round(float(np.mean(vals)), 5)


def test_score_batch_term_mapping_equals_the_per_query_form():
    """InvertedIndex.score_batch maps all term strings in one pass; the arrays must be the ones the per-query form
    (vocab.get per term, then engine.flatten_queries) hands to the same C-ABI call. No GPU: only the host mapping runs."""
    from improving_learned_index_b200 import engine
    from improving_learned_index_b200.inverted_index.inverted_index import InvertedIndex
    ix = InvertedIndex.__new__(InvertedIndex)
    ix.vocab = {f"t{i:03d}": i for i in range(500)}
    rng = np.random.default_rng(11)
    queries = []
    for qi in range(300):
        terms = [f"t{int(t):03d}" if t < 500 else f"unknown{int(t)}" for t in rng.integers(0, 560, size=int(rng.integers(0, 9)))]
        queries.append(terms if qi % 4 else tuple(terms))
    queries += [[], iter(["t001", "zzz", "t001"]), ["t499"]]           # empty, a one-shot iterator, a duplicate term
    want_flat, want_offs = engine.flatten_queries([ix._term_ids(q) for q in queries[:-3]] + [[], [1, -1, 1], [499]])
    flat, offs = ix._flat_term_ids(queries)
    assert flat.dtype == np.uint32 and offs.dtype == np.uint64
    assert np.array_equal(flat, want_flat) and np.array_equal(offs, want_offs)
    assert (flat == 0xFFFFFFFF).sum() > 0 and len(ix.vocab) == 500    # unknown terms became OOV, nothing was added to the vocabulary
    s = {"t002", "t003", "nope"}                                      # a set, as the reference's process_query returns
    f2, o2 = ix._flat_term_ids([s])
    assert sorted(f2.tolist()) == [2, 3, 0xFFFFFFFF] and o2.tolist() == [0, 3]
    ix.vocab["extra"] = 500                                           # a vocabulary that grew is noticed
    assert ix._flat_term_ids([["extra"]])[0].tolist() == [500]


# ------------------------------------------------------------------ host-side collection parser (no GPU involved)
def _same(a, b):
    return (a.vocab() == b.vocab() and np.array_equal(a.doc_offsets, b.doc_offsets)
            and np.array_equal(a.term_ids, b.term_ids) and np.array_equal(a.scores, b.scores))


def test_fast_parser_equals_reference_shaped_parser_on_golden(golden):
    from improving_learned_index_b200 import collection_io as C
    for name in ("kat", "small", "zeros"):
        g = golden(name)
        for lines in (g["lines"], g.get("quantized_lines", g["lines"])):
            text = ''.join(l + '\n' for l in lines)
            a, b = C.parse_bytes(text.encode(), C.DICT), C.parse_python(text, C.DICT)
            assert _same(a, b)
            if "quantized_lines" in g and lines is g["quantized_lines"]:
                assert a.vocab() == g["vocab"]
            if '' not in [l.strip() for l in lines]:
                assert _same(C.parse_bytes(text.encode(), C.SEQUENCE), C.parse_python(text, C.SEQUENCE))


@pytest.mark.parametrize("pieces", ["1", "3"])
def test_fast_parser_fuzz_against_python(monkeypatch, pieces):
    from improving_learned_index_b200 import collection_io as C
    monkeypatch.setenv("DI_B200_PARSE_THREADS", pieces)
    rng = np.random.default_rng(3)
    terms = ["a", "b c", "đá", "x|y", "t:1", "q,r", "naïve", "end ", "　lead", "tab\tin", "colon:", " sp"]
    nums = ["1", "0.5", "2.50", "1e3", "-3.25", "+7", ".5", "5.", "1E-2", "0", "inf", "-Infinity", "nan", "12.0", " 3 ", "4\t"]
    for trial in range(300):
        lines = []
        for _ in range(rng.integers(1, 6)):
            n = rng.integers(0, 5)
            pairs = [f"{terms[rng.integers(len(terms))]}: {nums[rng.integers(len(nums))]}" for _ in range(n)]
            pad = [" ", "", "\t", " ", "  "][rng.integers(5)]
            lines.append(pad + ', '.join(pairs) + pad)
        eol = ["\n", "\r\n", "\r"][rng.integers(3)]
        text = eol.join(lines) + (eol if rng.integers(2) else "")
        for mode in (C.DICT, C.SEQUENCE):
            try:
                want = C.parse_python(text, mode)
            except ValueError:
                with pytest.raises(ValueError):
                    C.parse_bytes(text.encode(), mode)
                continue
            got = C.parse_bytes(text.encode(), mode)
            assert got.vocab() == want.vocab() and np.array_equal(got.doc_offsets, want.doc_offsets), (trial, text)
            assert np.array_equal(got.term_ids, want.term_ids), (trial, text)
            assert np.array_equal(got.scores, want.scores, equal_nan=True), (trial, text)


def test_fast_parser_errors_and_fallback(tmp_path):
    from improving_learned_index_b200 import collection_io as C
    for bad in (b"a 1.5\n", b"a: 1: 2\n", b"a: x\n", b"a: 1,b: 2\n", b"a: \n"):
        with pytest.raises(ValueError):
            C.parse_bytes(bad, C.DICT)
        with pytest.raises(ValueError):
            C.parse_python(bad.decode(), C.DICT)
    with pytest.raises(ValueError):
        C.parse_bytes(b"a: 1\n\nb: 2\n", C.SEQUENCE)           # quantize.py:43 on a blank line
    assert C.parse_bytes(b"a: 1\n\nb: 2\n", C.DICT).doc_offsets.tolist() == [0, 1, 1, 2]
    with pytest.raises(_native.NativeError) as e:               # CPython accepts 1_0; the fast parser defers
        C.parse_bytes(b"a: 1_0\n", C.DICT)
    assert e.value.code == _native.ERR_UNSUPPORTED
    p = tmp_path / "c"
    p.write_bytes("a: 1_0, b: ٣\n".encode())                   # underscore literal, Arabic-Indic digit
    got = C.parse_file(p, C.DICT)
    assert got.scores.tolist() == [10.0, 3.0] and got.vocab() == ["a", "b"]
    (tmp_path / "bad").write_bytes(b"a: 1\n\xff\xfe: 2\n")
    with pytest.raises(UnicodeDecodeError):
        C.parse_file(tmp_path / "bad", C.DICT)
    # writer: quantize.py:40-47
    c = C.parse_bytes("x: 1, đ: 2\ny: 3\nz: 4\n".encode(), C.SEQUENCE)
    c.write_quantized(np.array([5, 0, -1, 9], dtype=np.int32), tmp_path / "out")
    assert (tmp_path / "out").read_text(encoding="utf-8") == "x: 5\n\nz: 9\n"


def test_fast_parser_is_independent_of_the_thread_count(monkeypatch, tmp_path):
    """The file is cut into one piece per host thread; arrays, vocabulary, the quantized text and the first
    error (with its line number) must not depend on how many pieces there are."""
    from improving_learned_index_b200 import collection_io as C
    from improving_learned_index_b200 import synthetic as syn
    c = syn.make_collection(400, vocab_size=300, draws_per_doc=25, seed=9)
    lines = c.lines()
    lines[17] = ""                                   # empty documents (DICT) ...
    lines[18] = ""
    lines[40] = lines[40] + ", " + lines[40]         # ... and a document that repeats its terms
    for eol in ("\n", "\r\n", "\r"):
        text = eol.join(lines) + eol
        monkeypatch.setenv("DI_B200_PARSE_THREADS", "1")
        base = C.parse_bytes(text.encode(), C.DICT)
        want = C.parse_python(text, C.DICT)
        assert _same(base, want)
        vals = (np.arange(base.term_ids.size) % 7).astype(np.int32)
        base.write_quantized(vals, tmp_path / "q1")
        for n in (2, 3, 7, 64, 256):
            monkeypatch.setenv("DI_B200_PARSE_THREADS", str(n))
            got = C.parse_bytes(text.encode(), C.DICT)
            assert _same(got, base), (eol, n)
            got.write_quantized(vals, tmp_path / "qn")
            assert (tmp_path / "qn").read_bytes() == (tmp_path / "q1").read_bytes()
    # first error in file order, same message whatever the cut
    bad = list(lines)
    bad[17] = bad[18] = "x: 1"
    bad[300] = "broken pair"
    bad[350] = "y: z"
    text = "\n".join(bad) + "\n"
    msgs = set()
    for n in (1, 2, 5, 64):
        monkeypatch.setenv("DI_B200_PARSE_THREADS", str(n))
        with pytest.raises(ValueError) as e:
            C.parse_bytes(text.encode(), C.SEQUENCE)
        msgs.add(str(e.value))
    assert len(msgs) == 1 and "line 301:" in msgs.pop()


def test_fast_parser_utf8_check_agrees_with_cpython(monkeypatch):
    """Invalid UTF-8 anywhere makes the fast parser defer (the caller's decode then raises UnicodeDecodeError)."""
    from improving_learned_index_b200 import collection_io as C
    good = ["é", "\u0800", "\ud7ff", "\ue000", "\uffff", "\U00010000", "\U0010ffff", "a" * 9 + "đ"]
    bad = [b"\x80", b"\xc0\xaf", b"\xc1\xbf", b"\xe0\x80\xaf", b"\xe0\x9f\xbf", b"\xed\xa0\x80", b"\xed\xbf\xbf",
           b"\xf0\x8f\xbf\xbf", b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"\xe2\x82", b"\xf0\x9f\x98", b"\xff"]
    for n in ("1", "4"):
        monkeypatch.setenv("DI_B200_PARSE_THREADS", n)
        for t in good:
            assert C.parse_bytes(f"{t}: 1\nb: 2\n".encode("utf-8"), C.DICT).vocab() == sorted([t, "b"])
        for raw in bad:
            for data in (raw + b": 1\nb: 2\n", b"b: 2\nlonger line here: 3\n" + raw + b": 1", b"a: 1, " + raw + b"\n"):
                with pytest.raises(UnicodeDecodeError):
                    data.decode("utf-8")
                with pytest.raises(_native.NativeError) as e:
                    C.parse_bytes(data, C.DICT)
                assert e.value.code == _native.ERR_UNSUPPORTED


def test_maxp_aggregation_matches_reference_output(golden, tmp_path):
    """aggregate_run (MaxP, aggregate_run.py:5-58): same rows as the reference script wrote for the same files."""
    from improving_learned_index_b200.aggregate_run import aggregate_run
    g = golden("maxp")
    (tmp_path / "run.tsv").write_text('\n'.join(g["run"]) + '\n', encoding='utf-8')
    (tmp_path / "mapping.txt").write_text('\n'.join(g["mapping"]) + '\n', encoding='utf-8')
    for top_k, want in g["out"].items():
        n = aggregate_run(tmp_path / "run.tsv", tmp_path / "mapping.txt", tmp_path / "out.tsv", int(top_k))
        got = (tmp_path / "out.tsv").read_text(encoding='utf-8').split('\n')[:-1]
        assert got == want and n == len(want)
    (tmp_path / "mixed.tsv").write_text("a\t0\t1\t2.0\n7\t0\t1\t1.0\n")
    with pytest.raises(TypeError):            # aggregate_run.py:52 cannot order 'a' against 7 either
        aggregate_run(tmp_path / "mixed.tsv", tmp_path / "mapping.txt", tmp_path / "o2.tsv")


def test_run_file_writer_matches_python_writelines(tmp_path, monkeypatch):
    """di_write_run_file (all host threads, pwrite at final offsets) appends exactly the bytes RunFile.writelines does
    (datasets.py:312-317), whatever the thread count: ragged counts, empty lists, non-ASCII ids, an existing file."""
    from improving_learned_index_b200.utils.datasets import RunFile
    rng = np.random.default_rng(3)
    n, k = 700, 90
    docids = rng.integers(0, 2 ** 32, size=(n, k), dtype=np.uint64).astype(np.uint32)
    scores = rng.integers(-5, 70000, size=(n, k)).astype(np.int32)
    counts = rng.integers(0, k + 1, size=n).astype(np.uint32)
    counts[:3] = [0, k, 1]
    qids = [f"q{i}" if i % 7 else f"ü{i}-äß" for i in range(n)]
    want = tmp_path / "python.tsv"
    want.write_text("0\t1\t1\t5\n", encoding="utf-8")
    ref = RunFile(want)
    for i, q in enumerate(qids):
        ref.writelines(q, list(zip(docids[i, :counts[i]].tolist(), scores[i, :counts[i]].tolist())))
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("DI_B200_IO_THREADS", threads)
        got = tmp_path / f"native{threads}.tsv"
        got.write_text("0\t1\t1\t5\n", encoding="utf-8")
        RunFile(got).write_batch(qids[:300], docids[:300], scores[:300], counts[:300])      # two appends, like two batches
        RunFile(got).write_batch(qids[300:], docids[300:], scores[300:], counts[300:])
        assert got.read_bytes() == want.read_bytes(), threads
        streamed = tmp_path / f"stream{threads}.tsv"
        streamed.write_text("0\t1\t1\t5\n", encoding="utf-8")
        with RunFile(streamed).stream() as out:                   # background copy of batch i while batch i+1 is formatted
            for lo in range(0, n, 150):
                d, s = docids[lo:lo + 150].copy(), scores[lo:lo + 150].copy()
                out.write_batch(qids[lo:lo + 150], d, s, counts[lo:lo + 150])
                d[:] = 0                                          # the arrays are the caller's again once write_batch returns
                s[:] = 0
        assert streamed.read_bytes() == want.read_bytes(), threads
    rows = list(RunFile(want).read())
    assert rows[1][2] == 1 and len(rows) == 1 + int(counts.sum())
    with pytest.raises(ValueError):
        RunFile(tmp_path / "x").write_batch(qids[:2], docids[:3], scores[:3], counts[:2])
