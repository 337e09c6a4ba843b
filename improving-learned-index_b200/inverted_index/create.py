"""InvertedIndexCreator — drop-in for src/deep_impact/inverted_index/create.py:12-68.

Same constructor, same ``run()``, same three output files, byte for byte:
``vocab.txt`` (terms in sorted() order), ``inverted_index.dat`` (5-byte postings, per term
by impact descending then docid ascending) and ``inverted_index.idx`` (byte ranges).
The term->document inversion itself (create.py:31-46: Python lists of tuples + sorted()) is
the GPU radix sort of csrc/build.cuh (K2); the 5-byte records are packed on the GPU too.
"""
from __future__ import annotations

import argparse
import struct
from pathlib import Path
from typing import Union

import numpy as np

from .. import collection_io, engine
from ..indexing.deep_impact_collection import DeepImpactCollection
from ..utils.defaults import (INVERTED_INDEX_DATA, INVERTED_INDEX_INDEX, INVERTED_INDEX_VOCAB, MAX_IMPACT)


class InvertedIndexCreator:
    def __init__(self, deep_impact_collection_path: Union[str, Path], output_path: Union[str, Path]):
        self._collection_path = Path(deep_impact_collection_path)
        self.deep_impact_collection = DeepImpactCollection(self._collection_path)
        self.output_path = Path(output_path)
        self.output_path.mkdir(parents=True, exist_ok=True)
        self.vocab = dict()
        self._docs = None

    def _parsed(self):
        if self._docs is None:
            self._docs = collection_io.parse_file(self._collection_path, collection_io.DICT)
        return self._docs

    def _vocab_file(self):
        """create.py:19-29 — term id = rank of the term in sorted() (code-point) order."""
        parsed = self._parsed()
        self.vocab = {term: i for i, term in enumerate(parsed.vocab())}
        (self.output_path / INVERTED_INDEX_VOCAB).write_bytes(parsed.vocab_file_bytes())

    def _inverted_index(self):
        parsed = self._parsed()
        scores = parsed.scores
        if scores.size and not np.isfinite(scores).all():       # int(nan) / int(inf) in create.py:35
            raise (ValueError('cannot convert float NaN to integer') if np.isnan(scores).any()
                   else OverflowError('cannot convert float infinity to integer'))
        values = np.trunc(scores)                               # create.py:35 stores int(val): truncation toward zero
        if values.size and (values.min() < 0 or values.max() > MAX_IMPACT):
            raise struct.error('ubyte format requires 0 <= number <= 255')     # what pack('B', val) raises
        if parsed.n_docs >= 2 ** 32:
            raise struct.error("'I' format requires 0 <= number <= 4294967295")
        toff, docids, impacts = engine.invert(parsed.term_ids, values.astype(np.uint8), parsed.doc_offsets,
                                              max(parsed.n_terms, 0))
        dat, idx = engine.serialize(toff, docids, impacts)
        dat.tofile(self.output_path / INVERTED_INDEX_DATA)
        idx.tofile(self.output_path / INVERTED_INDEX_INDEX)

    def run(self):
        self._vocab_file()
        self._inverted_index()
        if self._docs is not None:
            self._docs.close()
            self._docs = None


if __name__ == '__main__':
    args = argparse.ArgumentParser()
    args.add_argument('-i', '--deep_impact_collection_path', type=Path, required=True)
    args.add_argument('-o', '--output_path', type=Path, required=True)
    args = args.parse_args()
    InvertedIndexCreator(args.deep_impact_collection_path, args.output_path).run()
