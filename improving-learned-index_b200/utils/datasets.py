"""Wire formats either side of the search path: query files, qrels, run files.

Mirrors the parts of the reference's src/utils/datasets.py that the ranking path uses
(Queries :17-47, QueryRelevanceDataset :138-178, RunFile :305-324, QueryParser :370-389).
All ids are strings, as in the reference.
"""
from __future__ import annotations

import json
import logging
from pathlib import Path
from typing import Dict, Iterator, Optional, Set, Tuple, Union

from .defaults import COLLECTION_TYPES

PathLike = Union[str, Path]
logger = logging.getLogger(__name__)


class QueryParser:
    """One query line -> (qid, text). 'msmarco' = qid<TAB>text, 'beir' = JSON with _id/text."""

    @staticmethod
    def get_msmarco_item(query: str) -> Tuple[str, str]:
        qid, text = query.strip().split('\t')
        return str(qid), text

    @staticmethod
    def get_beir_item(query: str) -> Tuple[str, str]:
        record = json.loads(query)
        return record['_id'], record['text']

    @staticmethod
    def parse(item: str, collection_type: str) -> Tuple[str, str]:
        if collection_type == 'msmarco':
            return QueryParser.get_msmarco_item(item)
        if collection_type == 'beir':
            return QueryParser.get_beir_item(item)
        raise KeyError(collection_type)


class Queries:
    def __init__(self, queries_path: PathLike, dataset_type: Optional[str] = COLLECTION_TYPES[0]):
        self.dataset_type = dataset_type
        self.queries: Dict[str, str] = {}
        with open(queries_path, encoding='utf-8') as f:
            for line in f:
                qid, text = QueryParser.parse(line, dataset_type)
                self.queries[str(qid)] = text

    def __len__(self) -> int:
        return len(self.queries)

    def __getitem__(self, qid) -> str:
        return self.queries[str(qid)]

    def __iter__(self) -> Iterator[Tuple[str, str]]:
        return iter(self.queries.items())

    def keys(self):
        return self.queries.keys()


class QueryRelevanceDataset:
    """qrels: qid<TAB>0<TAB>pid<TAB>1 per line (any other 2nd/4th column is rejected, datasets.py:158)."""

    def __init__(self, qrels_path: PathLike):
        self.qrels: Dict[str, Set[str]] = {}
        with open(qrels_path, 'r', encoding='utf-8') as f:
            for line in f:
                cols = line.strip().split('\t')
                assert int(cols[1]) == 0 and int(cols[3]) == 1, "Qrels file is not in the expected format"
                self.qrels.setdefault(str(cols[0]), set()).add(str(cols[2]))
        # datasets.py:161-162: logged there; an empty qrels file is a ZeroDivisionError there and here
        self.average_positive_per_query = round(sum(len(p) for p in self.qrels.values()) / len(self.qrels), 2)
        logger.info(f"Loaded {len(self.qrels)} queries with {self.average_positive_per_query} positive passages/query on average")

    def __len__(self) -> int:
        return len(self.qrels)

    def __getitem__(self, qid) -> Set[str]:
        return self.qrels[str(qid)]

    def keys(self):
        return self.qrels.keys()


class RunFile:
    """qid<TAB>pid<TAB>rank<TAB>score rows, ranks from 1. Opens in APPEND mode per call, like the
    reference (datasets.py:310,315): re-running a ranker onto an existing file duplicates rows."""

    def __init__(self, run_file_path: PathLike):
        self.run_file_path = run_file_path

    def write(self, qid, pid, rank, score):
        with open(self.run_file_path, 'a', encoding='utf-8') as f:
            f.write(f'{qid}\t{pid}\t{rank}\t{score}\n')

    def writelines(self, qid, scores):
        rows = [f'{qid}\t{pid}\t{rank}\t{score}\n' for rank, (pid, score) in enumerate(scores, start=1)]
        with open(self.run_file_path, 'a', encoding='utf-8') as f:
            f.writelines(rows)

    @staticmethod
    def _batch_args(qids, docids, scores, counts):
        import numpy as np
        blob = [str(q).encode('utf-8') for q in qids]
        offs = np.zeros(len(blob) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(b) for b in blob])
        d = np.ascontiguousarray(docids, dtype=np.uint32)
        s = np.ascontiguousarray(scores, dtype=np.int32)
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        if d.ndim != 2 or d.shape != s.shape or d.shape[0] != len(blob) or c.shape != (len(blob),):
            raise ValueError("write_batch: docids/scores must be [len(qids), k] and counts [len(qids)]")
        return b''.join(blob), offs, d, s, c

    def write_batch(self, qids, docids, scores, counts):
        """writelines() for a whole batch straight from result arrays (docids / scores: [n, k] rows, counts[i] valid
        entries of row i): the same bytes, formatted and written by all host threads inside the library
        (di_write_run_file) instead of one Python string per row."""
        from .. import _native as N
        blob, offs, d, s, c = self._batch_args(qids, docids, scores, counts)
        N.check(N.lib().di_write_run_file(str(self.run_file_path).encode(), blob, N.ptr(offs), N.ptr(d), N.ptr(s),
                                          N.ptr(c), len(offs) - 1, d.shape[1]))

    def stream(self):
        """Context manager for many batches: `with run_file.stream() as out: out.write_batch(...)`. A batch is copied into
        the file in the background while the next one is searched and formatted (di_run_writer_*); the arrays passed
        to write_batch may be reused as soon as it returns."""
        return _RunStream(self)

    def read(self):
        with open(self.run_file_path, 'r', encoding='utf-8') as f:
            for line in f:
                qid, pid, rank, score = line.strip().split('\t')
                yield str(qid), str(pid), int(rank), float(score)


class _RunStream:
    def __init__(self, run_file: RunFile):
        self.run_file, self._w = run_file, None

    def __enter__(self):
        import ctypes
        from .. import _native as N
        self._w = ctypes.c_void_p()
        N.check(N.lib().di_run_writer_open(str(self.run_file.run_file_path).encode(), ctypes.byref(self._w)))
        return self

    def write_batch(self, qids, docids, scores, counts):
        from .. import _native as N
        blob, offs, d, s, c = RunFile._batch_args(qids, docids, scores, counts)
        N.check(N.lib().di_run_writer_submit(self._w, blob, N.ptr(offs), N.ptr(d), N.ptr(s), N.ptr(c), len(offs) - 1, d.shape[1]))

    def __exit__(self, exc_type, exc, tb):
        from .. import _native as N
        w, self._w = self._w, None
        rc = N.lib().di_run_writer_close(w)
        if exc_type is None:
            N.check(rc)
        return False
