"""Parser of the doc-major text collection: one document per line, ``term: score, term: score``;
the docid is the 0-based line number. Same behaviour as the reference's
src/deep_impact/indexing/deep_impact_collection.py:6-45 (blank line -> {}, a term repeated in
a line keeps its LAST score, separators are exactly ', ' and ': ').
"""
from __future__ import annotations

from itertools import permutations
from pathlib import Path
from typing import Dict, Set, Union


def parse_line(text: str) -> Dict[str, float]:
    """One stripped collection line -> {term: score} (insertion-ordered, last duplicate wins)."""
    if not text.strip():
        return {}
    impacts: Dict[str, float] = {}
    for pair in text.split(', '):
        term, score = pair.split(': ')          # ValueError on a malformed pair, as in the reference
        impacts[term] = float(score)
    return impacts


class DeepImpactCollection:
    def __init__(self, index_path: Union[str, Path]):
        with open(index_path, encoding='utf-8') as f:
            self.document_encodings = [line.strip() for line in f]

    def __len__(self) -> int:
        return len(self.document_encodings)

    def __getitem__(self, pid: int) -> Dict[str, float]:
        return parse_line(self.document_encodings[pid])

    def __iter__(self):
        for pid in range(len(self)):
            yield pid, self[pid]

    def score(self, pid: int, query_terms: Set[str]):
        impacts = self[pid]
        return sum(impacts.get(term, 0) for term in query_terms)


class DeepPairwiseImpactCollection(DeepImpactCollection):
    def score(self, pid: int, query_terms: Set[str]):
        impacts = self[pid]
        total = sum(impacts.get(term, 0) for term in query_terms)
        for a, b in permutations(query_terms, 2):
            total += impacts.get(f'{a}|{b}', 0)
        return total
