"""Ranker — drop-in for src/deep_impact/evaluation/ranker.py:14-58 (and the rank.py CLI).

Same constructor arguments and ``run()`` effect (rows ``qid<TAB>pid<TAB>rank<TAB>score`` appended
to the run file). The reference fans queries out over a multiprocessing.Pool and pickles the
index into every task (ranker.py:44-46); a GPU-resident index cannot be pickled, so queries are
scored in batches on the GPU instead and ``num_workers`` is accepted but unused.

Query text -> terms: the reference calls ``DeepImpactXLMR.process_query`` /
``DeepPairwiseImpact.process_query`` (tokenizers that need the HF hub). Pass any
``query_processor: str -> iterable of terms``; the default lower-cases and splits on whitespace.
"""
from __future__ import annotations

from itertools import product
from pathlib import Path
from typing import Callable, Iterable, Optional, Union

from ..inverted_index import InvertedIndex
from ..utils.datasets import Queries, QueryRelevanceDataset, RunFile
from ..utils.defaults import COLLECTION_TYPES


def whitespace_query_processor(query: str):
    return set(query.lower().split())


def rank(args):
    """Single-query form kept for API compatibility (ranker.py:14-16)."""
    index, qid, query_terms = args
    return qid, index.score(query_terms=query_terms)


class Ranker:
    def __init__(self, index_path: Union[str, Path], queries_path: Union[str, Path], output_path: Union[str, Path],
                 num_workers: int = 4, qrels_path: Optional[Union[str, Path]] = None, pairwise: bool = False,
                 dataset_type: Optional[str] = COLLECTION_TYPES[0],
                 query_processor: Optional[Callable[[str], Iterable[str]]] = None,
                 batch_size: int = 8192, top_k: int = 1000):
        self.queries = Queries(queries_path=queries_path, dataset_type=dataset_type)
        self.query_iterator = self.queries.keys()
        if qrels_path is not None:      # evaluate only the queries in the qrels file
            self.query_iterator = QueryRelevanceDataset(qrels_path=qrels_path).keys()
        self.index = InvertedIndex(index_path=index_path)
        self.run_file = RunFile(run_file_path=output_path)
        self.num_workers = num_workers
        self.pairwise = pairwise
        self.query_processor = query_processor or whitespace_query_processor
        self.batch_size = batch_size
        self.top_k = top_k

    def get_query_terms(self, qid):
        query_terms = set(self.query_processor(self.queries[qid]))
        if self.pairwise:
            for a, b in product(list(query_terms), repeat=2):
                if a != b:
                    query_terms.add(f'{a}|{b}')
        return query_terms

    def run(self):
        qids = list(self.query_iterator)
        for lo in range(0, len(qids), self.batch_size):
            batch = qids[lo:lo + self.batch_size]
            ranked = self.index.score_batch([self.get_query_terms(q) for q in batch], top_k=self.top_k)
            for qid, scores in zip(batch, ranked):
                self.run_file.writelines(qid, scores)
