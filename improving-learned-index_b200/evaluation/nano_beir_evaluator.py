"""SparseSearch / NanoBEIREvaluator — drop-in for the reference's
src/deep_impact/evaluation/nano_beir_evaluator.py (SparseSearch :70-137, BaseEvaluator
:139-151, NanoBEIREvaluator :153-232).

``SparseSearch`` keeps the reference's constructor, attributes and ``search`` signature and
return shape; the inversion of the model's (term, impact) lists and the query scoring run on
the GPU. The model is the same duck type: ``get_impact_scores_batch(list[str])`` and
``process_query(str)``. The GPU index is the 8-bit index the reference's own quantize step produces
(defaults.py:26): impacts must be integers in [0, 255], or ``quantize_max`` says how to get there —
a number: the reference rule int(v * 255 / quantize_max) (quantize.py:13-14,37); ``"auto"``: the same
rule with the collection's own maximum (find_max_value, quantize.py:17-24) whenever the model emits
non-integer impacts (what a real DeepImpact model does, models/original.py:309). ``NanoBEIREvaluator``
uses ``"auto"``. Scores are then sums of 8-bit impacts — the scale of the reference's on-disk index
path — not the float sums the reference's in-memory SparseSearch returns: documents whose float
scores differ by less than the quantization step may swap places (documented deviation).
Ties are ordered by corpus position (the reference: first-touch order, PYTHONHASHSEED-dependent).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Iterable, List, Optional

import numpy as np

from .. import engine
from .trec_metrics import EvaluateRetrieval

MAPPING_DATASET_NAME_TO_ID = {
    "climatefever": "zeta-alpha-ai/NanoClimateFEVER",
    "dbpedia": "zeta-alpha-ai/NanoDBPedia",
    "fever": "zeta-alpha-ai/NanoFEVER",
    "fiqa2018": "zeta-alpha-ai/NanoFiQA2018",
    "hotpotqa": "zeta-alpha-ai/NanoHotpotQA",
    "msmarco": "zeta-alpha-ai/NanoMSMARCO",
    "nfcorpus": "zeta-alpha-ai/NanoNFCorpus",
    "nq": "zeta-alpha-ai/NanoNQ",
    "quoraretrieval": "zeta-alpha-ai/NanoQuoraRetrieval",
    "scidocs": "zeta-alpha-ai/NanoSCIDOCS",
    "arguana": "zeta-alpha-ai/NanoArguAna",
    "scifact": "zeta-alpha-ai/NanoSciFact",
    "touche2020": "zeta-alpha-ai/NanoTouche2020",
}
MAPPING_DATASET_NAME_TO_HUMAN_READABLE = {
    name: repo.split("/Nano")[1] for name, repo in MAPPING_DATASET_NAME_TO_ID.items()
}


class Dataset:
    def __init__(self, queries, corpus, relevant_docs, name):
        self.queries = queries
        self.corpus = corpus
        self.relevant_docs = relevant_docs
        self.name = name


class SparseSearch:
    def __init__(self, model, batch_size, verbose=False, quantize_max: Optional[float] = None,
                 tile_docs: int = 0, dense_ratio: int = 0, cand_slack: int = 0):
        self.model = model
        self.batch_size = batch_size
        self.inverted_index = defaultdict(list)   # term -> [(doc_id, score), ...], as in the reference
        self.corpus_ids: List[str] = []
        self.verbose = verbose
        self.quantize_max = quantize_max
        self._index_params = dict(tile_docs=tile_docs, dense_ratio=dense_ratio, cand_slack=cand_slack)
        self._term_ids: Dict[str, int] = {}
        self.device_index: Optional[engine.DeviceIndex] = None

    def _build_inverted_index(self, corpus):
        """nano_beir_evaluator.py:78-101: run the model over the corpus in batches, keep postings
        with score > 0. Postings are collected doc-major and inverted on the GPU (K2)."""
        if self.verbose:
            print(f"Building inverted index for {len(corpus)} documents...")
        self.corpus_ids = list(corpus.keys())
        texts = list(corpus.values())
        term_ids, raw_scores, doc_offsets = [], [], [0]
        for lo in range(0, len(texts), self.batch_size):
            embeddings = self.model.get_impact_scores_batch(texts[lo:lo + self.batch_size])
            for doc_id, embedding in zip(self.corpus_ids[lo:lo + self.batch_size], embeddings):
                for term, score in embedding:
                    if score > 0:
                        self.inverted_index[term].append((doc_id, score))
                        term_ids.append(self._term_ids.setdefault(term, len(self._term_ids)))
                        raw_scores.append(float(score))
                doc_offsets.append(len(term_ids))
        scores = np.asarray(raw_scores, dtype=np.float64)
        quantize_max = self.quantize_max
        if isinstance(quantize_max, str):
            if quantize_max != "auto":
                raise ValueError("quantize_max must be a number, None or 'auto'")
            integral = scores.size == 0 or (np.array_equal(np.floor(scores), scores) and scores.max() <= 255)
            quantize_max = None if integral else engine.find_max(scores)
        if quantize_max is not None:
            values = engine.quantize(scores, quantize_max).astype(np.int64)
        else:
            values = scores.astype(np.int64)
            if scores.size and not np.array_equal(values, scores):
                raise ValueError("SparseSearch: impacts must be integers (8-bit quantized); pass quantize_max=<max impact> "
                                 "to quantize model outputs with the reference rule int(v * 255 / max)")
        if values.size and (values.min() < 0 or values.max() > 255):
            raise ValueError("SparseSearch: impacts must lie in [0, 255] after quantization")
        toff, docids, impacts = engine.invert(np.asarray(term_ids, dtype=np.uint32), values.astype(np.uint8),
                                              np.asarray(doc_offsets, dtype=np.uint64), max(len(self._term_ids), 1))
        self.device_index = engine.DeviceIndex.from_csr(toff, docids, impacts, **self._index_params)
        if self.verbose:
            print(f"Built inverted index with {len(self.inverted_index)} terms")

    def search(self, queries, corpus, k):
        if not self.inverted_index:
            self._build_inverted_index(corpus)
        if self.verbose:
            print(f"Searching for {len(queries)} queries...")
        query_ids = list(queries.keys())
        term_lists = [[self._term_ids.get(t, -1) for t in self.model.process_query(queries[q])] for q in query_ids]
        results = {q: {} for q in query_ids}
        n_docs = len(self.corpus_ids)
        if query_ids and n_docs and k > 0 and self.device_index is not None:
            docs, scores, counts = self.device_index.search(term_lists, min(int(k), n_docs))
            ids = self.corpus_ids
            for i, q in enumerate(query_ids):
                c = int(counts[i])
                results[q] = {ids[d]: float(s) for d, s in zip(docs[i, :c].tolist(), scores[i, :c].tolist())}
        if self.verbose:
            print(f"Retrieved top-{k} documents for {len(queries)} queries")
        return results


class BaseEvaluator:
    def __init__(self, batch_size=16, verbose=False):
        self.verbose = verbose
        self.batch_size = batch_size

    def _load_dataset(self, dataset_name) -> Dataset:
        pass

    def evaluate_dataset(self, model, dataset_name):
        pass

    def evaluate_all(self, model):
        pass


class NanoBEIREvaluator(BaseEvaluator):
    """``datasets`` (optional): {name: Dataset} to evaluate instead of downloading the 13 NanoBEIR
    sets from the HF hub (nano_beir_evaluator.py:165-167 needs the network)."""

    K_VALUES = [10, 100, 1000]

    def __init__(self, batch_size=16, verbose=False, datasets: Optional[Dict[str, Dataset]] = None, quantize_max="auto"):
        super().__init__(batch_size, verbose)
        self._datasets = datasets
        self.quantize_max = quantize_max   # how float model impacts reach the 8-bit index (see SparseSearch)

    def dataset_names(self) -> Iterable[str]:
        return list(self._datasets) if self._datasets is not None else list(MAPPING_DATASET_NAME_TO_ID)

    def _load_dataset(self, dataset_name) -> Dataset:
        if self._datasets is not None:
            return self._datasets[dataset_name]
        from datasets import load_dataset   # needs network access to the HF hub, like the reference
        repo = MAPPING_DATASET_NAME_TO_ID[dataset_name.lower()]
        corpus = load_dataset(repo, "corpus", split="train")
        queries = load_dataset(repo, "queries", split="train")
        qrels = load_dataset(repo, "qrels", split="train")
        relevant: Dict[str, Dict[str, int]] = {}
        for row in qrels:
            relevant.setdefault(row["query-id"], {})[row["corpus-id"]] = 1
        return Dataset(
            queries={r["_id"]: r["text"] for r in queries if len(r["text"]) > 0},
            corpus={r["_id"]: r["text"] for r in corpus if len(r["text"]) > 0},
            relevant_docs=relevant,
            name=MAPPING_DATASET_NAME_TO_HUMAN_READABLE[dataset_name],
        )

    def evaluate_dataset(self, model, dataset_name):
        dataset = self._load_dataset(dataset_name)
        searcher = SparseSearch(model, batch_size=self.batch_size, verbose=self.verbose, quantize_max=self.quantize_max)
        results = searcher.search(dataset.queries, dataset.corpus, k=1000)
        return EvaluateRetrieval().evaluate(dataset.relevant_docs, results, self.K_VALUES)

    def evaluate_all(self, model):
        metrics = {}
        for name in self.dataset_names():
            if self.verbose:
                print(f"Evaluating dataset {name}...")
            metrics[name] = self.evaluate_dataset(model, name)
            if self.verbose:
                print(f"Metrics for {name}: {metrics[name]}")
        names = list(metrics)
        # macro average per metric family, same tuple layout as one dataset's result
        metrics["avg"] = tuple(
            {f"{family}@{k}": sum(metrics[n][pos][f"{family}@{k}"] for n in names) / len(names) for k in self.K_VALUES}
            for pos, family in enumerate(("NDCG", "MAP", "Recall", "P")))
        return metrics
