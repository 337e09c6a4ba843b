"""CLI twin of the reference's src/deep_impact/rank.py: python -m improving_learned_index_b200.rank ..."""
import argparse
from pathlib import Path

from .evaluation import Ranker
from .utils.defaults import COLLECTION_TYPES

if __name__ == "__main__":
    parser = argparse.ArgumentParser("Rank queries against an inverted index on the GPU and write a run file.")
    parser.add_argument("--index_path", type=Path, required=True)
    parser.add_argument("--queries_path", type=Path, required=True)
    parser.add_argument("--output_path", type=Path, required=True)
    parser.add_argument("--num_workers", type=int, default=4, help="accepted for compatibility; unused")
    parser.add_argument("--qrels_path", type=Path, default=None)
    parser.add_argument("--dataset_type", type=str, default=COLLECTION_TYPES[0], choices=COLLECTION_TYPES)
    parser.add_argument("--pairwise", action='store_true')
    Ranker(**vars(parser.parse_args())).run()
