"""NDCG / MAP / Recall / P at cut-offs, with the call shape of
``beir.retrieval.evaluation.EvaluateRetrieval.evaluate`` that the reference uses at
src/deep_impact/evaluation/nano_beir_evaluator.py:230-231.

``beir`` and ``pytrec_eval`` are not vendored in the reference and not installed here, so this
restates trec_eval's published definitions. Against beir itself PARITY IS UNPINNED; the
arithmetic is checked against an independent published implementation, scikit-learn's ``ndcg_score`` /
``average_precision_score`` (tests/test_host_logic.py::test_trec_metrics_against_scikit_learn), and hand-computed
cases. The definitions:
ranking = score descending with ties broken by document id descending; ndcg_cut uses linear
gain and log2(rank + 1) discount with the ideal ranking taken from the qrels; map_cut sums
precision at relevant ranks <= k and divides by the number of relevant documents; recall.k and
P.k as usual. Values are averaged over the evaluated queries and rounded to 5 decimals (beir).
When ``beir`` is importable it is used instead.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple


def _rank(run: Dict[str, float]) -> List[str]:
    return [d for d, _ in sorted(run.items(), key=lambda kv: (kv[1], kv[0]), reverse=True)]


class EvaluateRetrieval:
    def evaluate(self, qrels: Dict[str, Dict[str, int]], results: Dict[str, Dict[str, float]],
                 k_values: List[int], ignore_identical_ids: bool = True
                 ) -> Tuple[Dict[str, float], Dict[str, float], Dict[str, float], Dict[str, float]]:
        try:  # pragma: no cover - exercised only where beir exists
            from beir.retrieval.evaluation import EvaluateRetrieval as _Beir
            return _Beir.evaluate(qrels, results, k_values, ignore_identical_ids)
        except ImportError:
            pass
        if ignore_identical_ids:
            results = {q: {d: s for d, s in docs.items() if d != q} for q, docs in results.items()}
        ndcg = {f"NDCG@{k}": 0.0 for k in k_values}
        _map = {f"MAP@{k}": 0.0 for k in k_values}
        recall = {f"Recall@{k}": 0.0 for k in k_values}
        precision = {f"P@{k}": 0.0 for k in k_values}
        evaluated = [q for q in qrels if q in results]
        for q in evaluated:
            rel = {d: g for d, g in qrels[q].items() if g > 0}
            ranking = _rank(results[q])
            gains = [rel.get(d, 0) for d in ranking]
            ideal = sorted(rel.values(), reverse=True)
            for k in k_values:
                top = gains[:k]
                dcg = sum(g / math.log2(i + 2) for i, g in enumerate(top))
                idcg = sum(g / math.log2(i + 2) for i, g in enumerate(ideal[:k]))
                ndcg[f"NDCG@{k}"] += dcg / idcg if idcg > 0 else 0.0
                hits, ap = 0, 0.0
                for i, g in enumerate(top):
                    if g > 0:
                        hits += 1
                        ap += hits / (i + 1)
                _map[f"MAP@{k}"] += ap / len(rel) if rel else 0.0
                recall[f"Recall@{k}"] += hits / len(rel) if rel else 0.0
                precision[f"P@{k}"] += hits / k
        n = max(len(evaluated), 1)
        for table in (ndcg, _map, recall, precision):
            for key in table:
                table[key] = round(table[key] / n, 5)
        return ndcg, _map, recall, precision
