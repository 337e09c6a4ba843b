"""Generate tests/golden/*.json by RUNNING THE REFERENCE'S OWN CODE in this container.

Run once here (``python oracle/make_golden.py``); the fixtures are committed because
/root/reference does not travel to the GPU box. The reference is copied to a temporary
directory first and imported from there — importing it in place would create ``logs/`` inside
/root/reference (src/utils/logger.py:20-21 runs at import of indexing/quantize.py:10).

What is recorded (all produced by unmodified reference functions):
  kat.json        the hand-checkable 4-document case of SURVEY.md §8c
  quantize.json   quantize() / find_max_value() / quantize_file() on edge values
  small.json      200-doc collection: quantized text, vocab, .idx/.dat bytes, raw and full score lists
  medium.json     3000-doc Zipf collection: file hashes + raw top-1000 lists for 30 queries
  zeros.json      collection holding literal 0 impacts: reader hides them (inverted_index.py:50-51)
  sparse.json     SparseSearch.search (nano_beir_evaluator.py:103-137) with a replaying fake model
  metrics.json    Metrics.evaluate sums (metrics.py:26-57) on a small run file + qrels
  maxp.json       aggregate_run.main (aggregate_run.py:5-58): passage run file -> MaxP document run file
                  (``python oracle/make_golden.py --only maxp`` regenerates just this one)
"""
from __future__ import annotations

import base64
import hashlib
import importlib
import importlib.util
import json
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
REFERENCE = Path("/root/reference")


def _load_pkg_synthetic():
    spec = importlib.util.spec_from_file_location(
        "di_synthetic", REPO / "improving-learned-index_b200" / "synthetic.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["di_synthetic"] = mod
    spec.loader.exec_module(mod)
    return mod


def _import_reference(tmp: Path):
    shutil.copytree(REFERENCE / "src", tmp / "src")
    sys.path.insert(0, str(tmp))
    # beir / datasets are not installed; SparseSearch itself needs neither.
    for name in ("beir", "beir.retrieval", "beir.retrieval.evaluation", "datasets"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["beir.retrieval.evaluation"].EvaluateRetrieval = object
    sys.modules["datasets"].load_dataset = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("offline"))
    quant = importlib.import_module("src.deep_impact.indexing.quantize")
    create = importlib.import_module("src.deep_impact.inverted_index.create")
    inv = importlib.import_module("src.deep_impact.inverted_index.inverted_index")
    coll = importlib.import_module("src.deep_impact.indexing.deep_impact_collection")

    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, tmp / rel)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    nano = by_path("ref_nano", "src/deep_impact/evaluation/nano_beir_evaluator.py")
    metrics = by_path("ref_metrics", "src/deep_impact/evaluation/metrics.py")
    return quant, create, inv, coll, nano, metrics


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode()


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def build_with_reference(ref, workdir: Path, lines, max_val=None, prequantized=False):
    quant, create, inv = ref[0], ref[1], ref[2]
    workdir.mkdir(parents=True, exist_ok=True)
    raw = workdir / "collection.index"
    raw.write_text(''.join(l + '\n' for l in lines), encoding='utf-8')
    if prequantized:
        qpath = raw
    else:
        qpath = workdir / "collection.index.quantized"
        quant.quantize_file(raw, qpath, max_val)
    out = workdir / "index"
    create.InvertedIndexCreator(qpath, out).run()
    return {
        "quantized_lines": qpath.read_text(encoding='utf-8').split('\n')[:-1],
        "vocab": (out / "vocab.txt").read_text(encoding='utf-8').split('\n')[:-1],
        "idx": (out / "inverted_index.idx").read_bytes(),
        "dat": (out / "inverted_index.dat").read_bytes(),
        "index": inv.InvertedIndex(out),
    }


def make_maxp(tmp: Path):
    """Run the reference's MaxP aggregation CLI (aggregate_run.py:5-58) on a small passage run file."""
    spec = importlib.util.spec_from_file_location("ref_aggregate_run", tmp / "src/deep_impact/aggregate_run.py")
    agg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(agg)
    wdir = tmp / "work" / "maxp"
    wdir.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(11)
    # passage ids "doc<d>#<p>", a few documents without '#', one id with two '#'
    mapping = [f"doc{d}#{p}" for d in range(12) for p in range(3)] + ["solo1", "solo2", "a#b#c"]
    rng.shuffle(mapping)
    run_rows = []
    for qid in ("10", "9", "100", "7"):
        pids = rng.choice(len(mapping) + 3, size=25, replace=False)          # some pids are not in the mapping
        scores = np.sort(rng.integers(0, 400, size=25))[::-1]
        for rank, (pid, sc) in enumerate(zip(pids.tolist(), scores.tolist()), start=1):
            run_rows.append(f"{qid}\t{pid}\t{rank}\t{sc}")
    run_rows.insert(5, "short\trow")                                            # skipped: fewer than 4 columns
    run_rows.append("7\t3\t26\t-4.5")                                           # a negative score never wins
    run_rows.append("11\t2\t1\t0")                                              # a query whose only score is 0
    (wdir / "run.tsv").write_text('\n'.join(run_rows) + '\n', encoding='utf-8')
    (wdir / "mapping.txt").write_text('\n'.join(mapping) + '\n', encoding='utf-8')
    outs = {}
    for top_k in (1000, 3):
        argv = sys.argv
        sys.argv = ["aggregate_run", "--run_file", str(wdir / "run.tsv"), "--mapping", str(wdir / "mapping.txt"),
                    "--output", str(wdir / f"out{top_k}.tsv"), "--top_k", str(top_k)]
        try:
            agg.main()
        finally:
            sys.argv = argv
        outs[str(top_k)] = (wdir / f"out{top_k}.tsv").read_text(encoding='utf-8').split('\n')[:-1]
    json.dump({"run": run_rows, "mapping": mapping, "out": outs}, open(GOLDEN / "maxp.json", "w"), indent=1)


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    syn = _load_pkg_synthetic()
    tmp = Path(tempfile.mkdtemp(prefix="di_ref_"))
    if sys.argv[1:3] == ["--only", "maxp"]:
        shutil.copytree(REFERENCE / "src", tmp / "src")
        make_maxp(tmp)
        shutil.rmtree(tmp)
        print("wrote", GOLDEN / "maxp.json")
        return
    ref = _import_reference(tmp)
    quant, create, inv, coll, nano, metrics = ref
    work = tmp / "work"

    # ---------------------------------------------------------------- kat.json
    kat_lines = ["apple: 2.5, banana: 0.004, cherry: 1.25", "banana: 5.0, apple: 2.5",
                 "cherry: 0.0, durian: 5.0, apple: 0.019", "apple: 2.5, cherry: 1.25, durian: 2.5"]
    b = build_with_reference(ref, work / "kat", kat_lines)
    index = b["index"]
    kat_queries = [(["apple", "cherry"], 3), (["apple", "apple"], 10), (["zzz"], 5),
                   (["durian", "banana", "zzz", "apple"], 2), ([], 4), (["cherry"], 1000)]
    json.dump({
        "lines": kat_lines, "max": quant.find_max_value(work / "kat" / "collection.index"),
        "quantized_lines": b["quantized_lines"], "vocab": b["vocab"], "idx": b64(b["idx"]), "dat": b64(b["dat"]),
        "term_docs": {t: index.term_docs(t) for t in ["apple", "banana", "cherry", "durian", "zzz"]},
        "term_location": {t: list(index.term_location(t)) for t in ["apple", "durian", "zzz"]},
        "scores": [{"terms": q, "top_k": k, "result": index.score(q, top_k=k)} for q, k in kat_queries],
    }, open(GOLDEN / "kat.json", "w"), indent=1)

    # ---------------------------------------------------------------- quantize.json
    rng = np.random.default_rng(11)
    maxima = [0.021, 0.042, 0.051, 0.055, 0.061, 0.084, 0.089, 0.091, 0.102, 0.105, 1.0, 3.217, 5.0,
              7.001, 11.999, 12.0, 19.731, 255.0, 1e-3, 123456.789]
    cases = []
    for m in maxima:
        vals = sorted(set([m, 0.0, 0.001, m / 2, m / 3, m / 255, m / 255 * 2, m * 254 / 255, m * 0.999]
                          + [round(float(x), 3) for x in rng.uniform(0, m, 12)]))
        scale = ((1 << 8) - 1) / m
        cases.append({"max": m, "values": vals, "quantized": [quant.quantize(v, scale) for v in vals]})
    sweep = [i / 1000 for i in range(1, 20001)]
    sweep_q = [quant.quantize(v, 255 / v) for v in sweep]            # self-max: 254 or 255
    qf_lines = ["a: 0.5, b: 1.0, c: 0.003", "d: 0.002", "a: 1.0, e: 0.999, f: 0.0039, g: 0.004", "b: 0.75"]
    (work / "qf").mkdir(parents=True)
    (work / "qf" / "in").write_text(''.join(l + '\n' for l in qf_lines))
    quant.quantize_file(work / "qf" / "in", work / "qf" / "out_auto")
    quant.quantize_file(work / "qf" / "in", work / "qf" / "out_m2", 2.0)
    quant.quantize_file(work / "qf" / "in", work / "qf" / "out_m05", 0.5)
    json.dump({
        "cases": cases,
        "sweep_254": [i + 1 for i, q in enumerate(sweep_q) if q == 254],  # m (in 1/1000) whose own max lands on 254
        "sweep_other": [i + 1 for i, q in enumerate(sweep_q) if q not in (254, 255)],
        "file": {"lines": qf_lines,
                 "auto": (work / "qf" / "out_auto").read_text().split('\n')[:-1],
                 "max2": (work / "qf" / "out_m2").read_text().split('\n')[:-1],
                 "max05": (work / "qf" / "out_m05").read_text().split('\n')[:-1]},
    }, open(GOLDEN / "quantize.json", "w"))

    # ---------------------------------------------------------------- small.json
    c = syn.make_collection(200, vocab_size=60, draws_per_doc=14, seed=3, zero_frac=0.05)
    lines = c.lines()
    b = build_with_reference(ref, work / "small", lines)
    index = b["index"]
    qs = syn.make_queries(24, vocab_size=60, seed=5, mean_extra=2.0, max_len=8)
    qterms = [[syn.term_name(t) for t in q] for q in qs]
    qterms[3] = qterms[3] + ["not-in-vocab"]
    qterms[4] = []
    qterms[5] = qterms[5] + qterms[5][:1]           # duplicated term counts twice (list input)
    json.dump({
        "lines": lines, "quantized_lines": b["quantized_lines"], "vocab": b["vocab"],
        "idx": b64(b["idx"]), "dat": b64(b["dat"]), "n_docs": 200,
        "queries": [{"terms": q, "all": index.score(q, top_k=10 ** 9), "top10": index.score(q, top_k=10),
                     "top1": index.score(q, top_k=1)} for q in qterms],
    }, open(GOLDEN / "small.json", "w"))

    # ---------------------------------------------------------------- medium.json
    c = syn.make_collection(3000, vocab_size=2000, draws_per_doc=60, seed=1)
    lines = c.lines()
    b = build_with_reference(ref, work / "medium", lines)
    index = b["index"]
    qs = syn.make_queries(30, vocab_size=2000, seed=9)
    qterms = [[syn.term_name(t) for t in q] for q in qs]
    json.dump({
        "gen": {"n_docs": 3000, "vocab_size": 2000, "draws_per_doc": 60, "seed": 1,
                "q": {"n": 30, "vocab_size": 2000, "seed": 9}},
        "lines_sha256": sha(''.join(l + '\n' for l in lines).encode()),
        "quantized_sha256": sha(''.join(l + '\n' for l in b["quantized_lines"]).encode()),
        "vocab_sha256": sha(''.join(t + '\n' for t in b["vocab"]).encode()),
        "idx_sha256": sha(b["idx"]), "dat_sha256": sha(b["dat"]),
        "n_postings": len(b["dat"]) // 5,
        "queries": [{"terms": q, "top1000": index.score(q, top_k=1000),
                     "n_touched": len(index.score(q, top_k=10 ** 9))} for q in qterms],
    }, open(GOLDEN / "medium.json", "w"))

    # ---------------------------------------------------------------- zeros.json
    zero_lines = ["x: 3, y: 0, z: 7", "x: 0, z: 7", "", "y: 2, x: 3, w: 0", "z: 1, y: 0"]
    b = build_with_reference(ref, work / "zeros", zero_lines, prequantized=True)
    index = b["index"]
    json.dump({
        "lines": zero_lines, "vocab": b["vocab"], "idx": b64(b["idx"]), "dat": b64(b["dat"]),
        "term_docs": {t: index.term_docs(t) for t in ["w", "x", "y", "z"]},
        "scores": [{"terms": q, "result": index.score(q, top_k=10)}
                   for q in (["x", "y"], ["w"], ["z", "y", "x", "w"])],
    }, open(GOLDEN / "zeros.json", "w"), indent=1)

    # ---------------------------------------------------------------- sparse.json
    c = syn.make_collection(300, vocab_size=80, draws_per_doc=16, seed=21, zero_frac=0.04)
    q8 = np.minimum(255, np.floor(c.impacts * 40)).astype(np.int64)     # integer impacts 0..255
    corpus, replay = {}, {}
    for d in range(c.n_docs):
        lo, hi = int(c.doc_offsets[d]), int(c.doc_offsets[d + 1])
        text = f"doc text {d}"
        corpus[f"D{d:04d}"] = text
        replay[text] = [(syn.term_name(int(t)), np.float32(v)) for t, v in zip(c.term_ids[lo:hi], q8[lo:hi])]
    qs = syn.make_queries(12, vocab_size=80, seed=23, mean_extra=2.0, max_len=6)
    queries = {f"Q{i}": ' '.join(syn.term_name(t) for t in q) for i, q in enumerate(qs)}
    queries["Qoov"] = "nothing matches"
    queries["Qmix"] = queries["Q0"] + " unknownterm"

    class FakeModel:
        def get_impact_scores_batch(self, texts):
            return [replay[t] for t in texts]

        def process_query(self, query):
            return list(dict.fromkeys(query.split()))   # ordered, so the raw tie order is reproducible
    res10 = nano.SparseSearch(FakeModel(), batch_size=16).search(queries, corpus, k=10)
    res_all = nano.SparseSearch(FakeModel(), batch_size=7).search(queries, corpus, k=1000)
    json.dump({
        "corpus": corpus,
        "replay": {k: [(t, float(v)) for t, v in lst] for k, lst in replay.items()},
        "queries": queries,
        "k10": {q: list(r.items()) for q, r in res10.items()},
        "k1000": {q: list(r.items()) for q, r in res_all.items()},
    }, open(GOLDEN / "sparse.json", "w"))

    # ---------------------------------------------------------------- metrics.json
    mdir = work / "metrics"
    mdir.mkdir(parents=True)
    rng = np.random.default_rng(5)
    run_rows, qrels_rows = [], []
    for qi in range(12):
        pids = rng.permutation(60)[:25]
        for rank, pid in enumerate(pids, start=1):
            run_rows.append(f"q{qi}\t{pid}\t{rank}\t{100 - rank}")
        rel = list(pids[rng.permutation(25)[: 1 + qi % 3]])          # ranked relevant docs
        if qi % 4 == 0:
            rel.append(int(rng.integers(60, 70)))                    # plus one never retrieved
        for pid in rel:
            qrels_rows.append(f"q{qi}\t0\t{pid}\t1")
    qrels_rows.append("q_unranked\t0\t5\t1")
    (mdir / "run.tsv").write_text('\n'.join(run_rows) + '\n')
    (mdir / "qrels.tsv").write_text('\n'.join(qrels_rows) + '\n')
    m = metrics.Metrics(mdir / "run.tsv", mdir / "qrels.tsv", mrr_depths=[10, 100], recall_depths=[3, 10, 20, 50])
    m.evaluate()
    n_q = len(m.qrels)
    json.dump({
        "run": run_rows, "qrels": qrels_rows, "n_qrels_queries": n_q,
        "mrr": {str(d): round(m.mrr_sums[d] / n_q, 3) for d in m.mrr_sums},
        "recall": {str(d): round(m.recall_sums[d] / n_q, 3) for d in m.recall_sums},
    }, open(GOLDEN / "metrics.json", "w"), indent=1)

    make_maxp(tmp)
    shutil.rmtree(tmp)
    print("golden fixtures written to", GOLDEN)
    for f in sorted(GOLDEN.iterdir()):
        print(f"  {f.name:16s} {f.stat().st_size:>9d} B")


if __name__ == "__main__":
    main()
