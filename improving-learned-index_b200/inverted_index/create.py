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

from .. import engine
from ..indexing.deep_impact_collection import DeepImpactCollection
from ..utils.defaults import (INVERTED_INDEX_DATA, INVERTED_INDEX_INDEX, INVERTED_INDEX_VOCAB, MAX_IMPACT)


class InvertedIndexCreator:
    def __init__(self, deep_impact_collection_path: Union[str, Path], output_path: Union[str, Path]):
        self.deep_impact_collection = DeepImpactCollection(Path(deep_impact_collection_path))
        self.output_path = Path(output_path)
        self.output_path.mkdir(parents=True, exist_ok=True)
        self.vocab = dict()
        self._docs = None

    def _parsed(self):
        if self._docs is None:
            self._docs = [item for _, item in self.deep_impact_collection]
        return self._docs

    def _vocab_file(self):
        """create.py:19-29 — term id = rank of the term in sorted() (code-point) order."""
        terms = set()
        for item in self._parsed():
            terms.update(item.keys())
        self.vocab = {term: i for i, term in enumerate(sorted(terms))}
        with open(self.output_path / INVERTED_INDEX_VOCAB, 'w', encoding='utf-8') as f:
            f.writelines(f'{term}\n' for term in self.vocab)

    def _inverted_index(self):
        docs = self._parsed()
        offsets = np.zeros(len(docs) + 1, dtype=np.uint64)
        np.cumsum([len(d) for d in docs], out=offsets[1:])
        n_post = int(offsets[-1])
        term_ids = np.fromiter((self.vocab[t] for d in docs for t in d), dtype=np.uint32, count=n_post)
        # create.py:35 stores int(val): truncation toward zero of the parsed float
        values = np.fromiter((int(v) for d in docs for v in d.values()), dtype=np.int64, count=n_post)
        if n_post and (values.min() < 0 or values.max() > MAX_IMPACT):
            raise struct.error('ubyte format requires 0 <= number <= 255')     # what pack('B', val) raises
        if len(docs) >= 2 ** 32:
            raise struct.error("'I' format requires 0 <= number <= 4294967295")
        toff, docids, impacts = engine.invert(term_ids, values.astype(np.uint8), offsets, len(self.vocab))
        dat, idx = engine.serialize(toff, docids, impacts)
        dat.tofile(self.output_path / INVERTED_INDEX_DATA)
        idx.tofile(self.output_path / INVERTED_INDEX_INDEX)

    def run(self):
        self._vocab_file()
        self._inverted_index()


if __name__ == '__main__':
    args = argparse.ArgumentParser()
    args.add_argument('-i', '--deep_impact_collection_path', type=Path, required=True)
    args.add_argument('-o', '--output_path', type=Path, required=True)
    args = args.parse_args()
    InvertedIndexCreator(args.deep_impact_collection_path, args.output_path).run()
