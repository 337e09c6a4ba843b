"""MRR@k / Recall@k over a run file and qrels — same numbers as the reference's
src/deep_impact/evaluation/metrics.py:13-74 (note its convention: the sums only receive
queries with at least one retrieved relevant passage, the divisor is ALL qrels queries, and
results are rounded to 3 decimals). CPU work; it consumes what the GPU ranker wrote.
"""
from __future__ import annotations

import logging
from collections import defaultdict
from pathlib import Path
from typing import Dict, List, Union

from ..utils.datasets import QueryRelevanceDataset, RunFile

logger = logging.getLogger('metrics')


class Metrics:
    def __init__(self, run_file_path: Union[str, Path], qrels_path: Union[str, Path],
                 mrr_depths: List[int], recall_depths: List[int]):
        self.run_file = RunFile(run_file_path=run_file_path)
        self.qrels = QueryRelevanceDataset(qrels_path=qrels_path)
        self.mrr_sums = {depth: 0 for depth in mrr_depths}
        self.recall_sums = {depth: 0 for depth in recall_depths}

    def evaluate(self) -> Dict[str, float]:
        """Logs MRR@d and Recall@d like the reference; additionally returns them as a dict."""
        hit_ranks = defaultdict(list)                      # qid -> ranks of its retrieved relevant pids
        for qid, pid, rank, _ in self.run_file.read():
            if pid in self.qrels[qid]:
                hit_ranks[qid].append(rank)

        for qid, ranks in hit_ranks.items():
            best = min(ranks)
            for depth in self.mrr_sums:
                if best <= depth:
                    self.mrr_sums[depth] += 1.0 / best
            n_rel = len(self.qrels[qid])
            for depth in self.recall_sums:
                self.recall_sums[depth] += sum(1 for r in ranks if r <= depth) / n_rel

        n_queries = len(self.qrels)
        logger.info(f"\nEvaluated {n_queries} queries")
        report = {}
        for depth in sorted(self.mrr_sums):
            report[f'MRR@{depth}'] = round(self.mrr_sums[depth] / n_queries, 3)
            logger.info(f"MRR@{depth} = {report[f'MRR@{depth}']}")
        for depth in sorted(self.recall_sums):
            report[f'Recall@{depth}'] = round(self.recall_sums[depth] / n_queries, 3)
            logger.info(f"Recall@{depth} = {report[f'Recall@{depth}']}")
        return report

    @staticmethod
    def evaluate_recall_for_top_k(qrels: QueryRelevanceDataset, top_k) -> float:
        """Recall at maximum depth of a top-k dataset (object with .queries, .keys(), [qid], .max_len)."""
        assert set(top_k.queries.keys()).issubset(set(qrels.keys())), "TopK file contains queries not in the Qrels file"
        per_query = [len(qrels[qid].intersection(set(top_k[qid]))) / len(qrels[qid]) for qid in top_k.keys()]
        recall = round(sum(per_query) / len(per_query), 3)
        logger.info(f"Recall@{top_k.max_len} = {recall}")
        return recall
