"""CLI twin of the reference's src/deep_impact/evaluate.py."""
import argparse
from pathlib import Path

from .evaluation import Metrics

MRR_DEPTHS = [10]
RECALL_DEPTHS = [3, 10, 20, 50] + list(range(100, 1001, 100))

if __name__ == "__main__":
    parser = argparse.ArgumentParser("Compute MRR / Recall of a run file against qrels.")
    parser.add_argument("--run_file_path", type=Path, required=True)
    parser.add_argument("--qrels_path", type=Path, required=True)
    args = parser.parse_args()
    print(Metrics(**vars(args), mrr_depths=MRR_DEPTHS, recall_depths=RECALL_DEPTHS).evaluate())
