"""B200-native inverted-index search for DeeperImpact (Tommachilez/improving-learned-index).

One hot path, rebuilt for sm_100a behind the reference's own Python surface: impact
quantization, term->document inversion, batched term-at-a-time scoring, deterministic top-k
and a docid-range sharded multi-GPU merge. See DESIGN.md / INTEGRATION.md at the repo root.

Importing this package does not need a GPU; any compute call does, and raises without one
(there is no CPU fallback).
"""
from . import engine
from .engine import DeviceIndex
from .indexing.deep_impact_collection import DeepImpactCollection, DeepPairwiseImpactCollection
from .indexing.quantize import find_max_value, quantize, quantize_file
from .inverted_index import InvertedIndex, InvertedIndexCreator

__all__ = [
    'engine', 'DeviceIndex', 'DeepImpactCollection', 'DeepPairwiseImpactCollection',
    'find_max_value', 'quantize', 'quantize_file', 'InvertedIndex', 'InvertedIndexCreator',
]
