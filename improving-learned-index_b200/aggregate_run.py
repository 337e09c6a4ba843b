"""MaxP aggregation of a passage-level run file into a document-level run file.

Host-side twin of the reference's `python -m src.deep_impact.aggregate_run` (aggregate_run.py:5-58), the
post-processing step after `rank` when documents were indexed as sliding-window passages: a passage id is the
0-based line number of `--mapping`, whose line holds the real id `doc#passage`; a document's score for a query
is the best score of its passages. Same flags, same output text (`qid \\t doc \\t rank \\t score:.6f`).

Behaviour kept from the reference, on purpose:
  * rows with fewer than 4 columns and passage ids missing from the mapping are skipped (:29, :36-37);
  * the document id is the text before the FIRST '#' (:41-44);
  * a document enters a query's list as soon as one of its passages is seen, with score 0.0 until a passage
    beats that (:47: the comparison itself creates the entry), so scores <= 0 are reported as 0.000000;
  * queries are written in numeric order when every id is a digit string (:52; a mix of digit and non-digit
    ids raises TypeError there and here), documents by descending score, ties in first-seen order (:54).
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Dict, Union


def aggregate_run(run_file: Union[str, Path], mapping: Union[str, Path], output: Union[str, Path], top_k: int = 1000) -> int:
    """Returns the number of rows written."""
    with open(mapping, 'r', encoding='utf-8') as f:
        real_id = {str(i): line.strip() for i, line in enumerate(f)}
    best: Dict[str, Dict[str, float]] = {}
    with open(run_file, 'r', encoding='utf-8') as f:
        for line in f:
            cols = line.strip().split('\t')
            if len(cols) < 4:
                continue
            passage = real_id.get(cols[1])
            if passage is None:
                continue
            score = float(cols[3])
            doc = passage.split('#', 1)[0]
            per_query = best.setdefault(cols[0], {})
            if score > per_query.setdefault(doc, 0.0):
                per_query[doc] = score
    written = 0
    with open(output, 'w', encoding='utf-8') as f:
        for qid in sorted(best, key=lambda x: int(x) if x.isdigit() else x):
            ranked = sorted(best[qid].items(), key=lambda kv: kv[1], reverse=True)[:top_k]
            for rank, (doc, score) in enumerate(ranked, start=1):
                f.write(f"{qid}\t{doc}\t{rank}\t{score:.6f}\n")
            written += len(ranked)
    return written


def main():
    parser = argparse.ArgumentParser(description="MaxP: passage run file -> document run file")
    parser.add_argument("--run_file", required=True, help="run file written by rank (integer passage ids)")
    parser.add_argument("--mapping", required=True, help="pid mapping: line i = real id of passage i")
    parser.add_argument("--output", required=True, help="document-level run file")
    parser.add_argument("--top_k", type=int, default=1000)
    args = parser.parse_args()
    n = aggregate_run(args.run_file, args.mapping, args.output, args.top_k)
    print(f"wrote {n} rows to {args.output}")


if __name__ == "__main__":
    main()
