"""Ranker — drop-in for src/deep_impact/evaluation/ranker.py:14-58 (and the rank.py CLI).

Same constructor arguments and ``run()`` effect (rows ``qid<TAB>pid<TAB>rank<TAB>score`` appended
to the run file). The reference fans queries out over a multiprocessing.Pool and pickles the
index into every task (ranker.py:44-46); a GPU-resident index cannot be pickled, so queries are
scored in batches on the GPU instead and ``num_workers`` is accepted but unused.

``run()`` hands whole batches to the library's run-file writer (all host threads, di_write_run_file) while the GPU
already scores the next batch, and — when it is given depths and the Ranker has qrels — also returns the MRR@k /
Recall@k report of ``Metrics.evaluate`` computed from the result keys while they are still on the device
(di_eval_ranks_dev), so the run file need not be parsed back.

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
        self.qrels = None
        if qrels_path is not None:      # evaluate only the queries in the qrels file
            self.qrels = QueryRelevanceDataset(qrels_path=qrels_path)
            self.query_iterator = self.qrels.keys()
        # (an already loaded InvertedIndex is accepted in place of its path: loading puts the whole index into HBM)
        self.index = index_path if isinstance(index_path, InvertedIndex) else InvertedIndex(index_path=index_path)
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

    def run(self, mrr_depths=None, recall_depths=None):
        """Writes the run file like the reference's Ranker.run. With depths (and a qrels_path given to the
        constructor) it also returns the report Metrics(run_file, qrels, mrr_depths, recall_depths).evaluate() would
        compute from that file — same numbers, taken from the device-resident results."""
        from concurrent.futures import ThreadPoolExecutor
        qids = list(self.query_iterator)
        want_metrics = (mrr_depths or recall_depths) and self.qrels is not None
        collector = _DeviceMetrics(self.qrels, mrr_depths or [], recall_depths or []) if want_metrics else None
        batches = [qids[lo:lo + self.batch_size] for lo in range(0, len(qids), self.batch_size)]

        def prepare(batch):                                     # query text -> term strings -> term ids (pure Python)
            return [self.index._term_ids(self.get_query_terms(q)) for q in batch]

        # four things overlap: Python prepares batch i+1, the GPU scores batch i, the library formats batch i-1 (writer
        # thread) and copies batch i-2 into the file (its own background threads); ctypes releases the GIL inside the
        # library, single workers keep the order
        with ThreadPoolExecutor(max_workers=1) as prep, ThreadPoolExecutor(max_workers=1) as writer, \
                self.run_file.stream() as out:
            ready = prep.submit(prepare, batches[0]) if batches else None
            pending = None
            for i, batch in enumerate(batches):
                term_ids = ready.result()
                if i + 1 < len(batches):
                    ready = prep.submit(prepare, batches[i + 1])
                if collector is not None:
                    docs, scores, counts = collector.search(self.index, batch, term_ids, self.top_k)
                else:
                    res = self.index.score_id_batch(term_ids, top_k=self.top_k, pinned=True)
                    docs, scores, counts = res.docids, res.scores, res.counts
                if pending is not None:
                    pending.result()
                pending = writer.submit(out.write_batch, batch, docs, scores, counts)
            if pending is not None:
                pending.result()
        return collector.report() if collector is not None else None


class _DeviceMetrics:
    """metrics.py:26-57 without the run file: per batch the result keys stay on the device, one kernel finds every
    query's first relevant rank and its relevant hits within each depth; the float sums are then formed on the host
    in run-file order exactly as Metrics.evaluate forms them."""

    def __init__(self, qrels, mrr_depths, recall_depths):
        self.qrels = qrels
        self.mrr_depths, self.recall_depths = list(mrr_depths), list(recall_depths)
        self.mrr_sums = {d: 0 for d in self.mrr_depths}
        self.recall_sums = {d: 0 for d in self.recall_depths}

    def _relevant_docids(self, qid):
        try:
            pids = self.qrels[qid]
        except KeyError:
            return []
        # the run file holds str(docid): only a canonical decimal pid can ever equal it (metrics.py:32 compares strings)
        return sorted(int(p) for p in pids if p.isdigit() and str(int(p)) == p and int(p) < 2 ** 32)

    def search(self, index, qids, term_lists, top_k):
        import numpy as np
        import torch
        from .. import _native as N, engine
        dev = torch.device("cuda", torch.cuda.current_device())
        st = torch.cuda.current_stream().cuda_stream
        n = len(qids)
        k = min(int(top_k), max(int(index._n_docs_hint), 1))
        flat, offs = engine.flatten_queries(term_lists)          # term ids already
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint32)
        d_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).to(dev)
        d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
        keys = torch.zeros((n, k), dtype=torch.int64, device=dev)
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        index.device_index.search_device(d_flat, d_offs, n, max((len(t) for t in term_lists), default=0), k, keys, counts, st)
        rel = [self._relevant_docids(q) for q in qids]
        r_offs = np.zeros(n + 1, dtype=np.int64)
        r_offs[1:] = np.cumsum([len(r) for r in rel])
        r_docs = np.asarray([d for r in rel for d in r] or [0], dtype=np.int64)
        depths = sorted(set(self.recall_depths)) or [1]
        d_roffs = torch.from_numpy(r_offs).to(dev)                      # named: they must outlive the kernel launches
        d_rdocs = torch.from_numpy(r_docs).to(torch.int32).to(dev)
        best = torch.zeros(n, dtype=torch.int32, device=dev)
        hits = torch.zeros((n, len(depths)), dtype=torch.int32, device=dev)
        parts = []
        for j0 in range(0, len(depths), 8):                             # the kernel takes up to 8 depths per launch
            d_part = torch.tensor(depths[j0:j0 + 8], dtype=torch.int32, device=dev)
            h = torch.zeros((n, d_part.numel()), dtype=torch.int32, device=dev)
            N.check(N.lib().di_eval_ranks_dev(keys.data_ptr(), counts.data_ptr(), n, k, d_roffs.data_ptr(), d_rdocs.data_ptr(),
                                              d_part.data_ptr(), d_part.numel(), best.data_ptr(), h.data_ptr(), st))
            parts.append((j0, d_part, h))
        for j0, d_part, h in parts:
            hits[:, j0:j0 + d_part.numel()] = h
        docs = torch.zeros((n, k), dtype=torch.int32, device=dev)
        scores = torch.zeros((n, k), dtype=torch.int32, device=dev)
        engine.unpack_keys_device(keys, n * k, docs, scores, st)
        torch.cuda.synchronize()
        best_h, hits_h = best.cpu().numpy(), hits.cpu().numpy()
        for i, qid in enumerate(qids):                       # run-file order, metrics.py:35-43
            if best_h[i] == 0:
                continue
            b = int(best_h[i])
            for depth in self.mrr_sums:
                if b <= depth:
                    self.mrr_sums[depth] += 1.0 / b
            n_rel = len(self.qrels[qid])
            for depth in self.recall_sums:
                self.recall_sums[depth] += int(hits_h[i, depths.index(depth)]) / n_rel
        return docs.cpu().numpy().view(np.uint32), scores.cpu().numpy(), counts.cpu().numpy().view(np.uint32)

    def report(self):
        n_queries = len(self.qrels)
        out = {f'MRR@{d}': round(self.mrr_sums[d] / n_queries, 3) for d in sorted(self.mrr_sums)}
        out.update({f'Recall@{d}': round(self.recall_sums[d] / n_queries, 3) for d in sorted(self.recall_sums)})
        return out
