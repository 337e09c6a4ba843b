#!/usr/bin/env python
"""Times the reference's OWN, unmodified Python reader on index files this repo's builder wrote.

    python tools/py_reference_timing.py --index-dir DIR [--queries 6980] [--seconds 12]

Imports `InvertedIndex` (src/deep_impact/inverted_index/inverted_index.py:19-62) from the copy of the reference
under baseline/_ref/ (made by __graft_entry__.build(); /root/reference itself does not exist on the GPU box) and
runs it on a bounded sample of the benchmark's query set, both ways the reference itself runs it:
  (i)  single process: a loop of `index.score(terms, 1000)`;
  (ii) `multiprocessing.Pool(cores).imap_unordered(rank, ...)` with `rank(args)` = `index.score(query_terms)`, exactly
       the fan-out of Ranker.run (evaluation/ranker.py:14-16, 44-46; the Ranker class itself cannot be imported
       offline because its module pulls the HF tokenizers).
The reference reads ~0.7 M postings/s per core, i.e. ~30 s for an average query of this index, so the sample is the
random queries (seed 123) whose posting count fits the time budget; the cost of the reference is linear in postings,
so postings/s is measured and the full-query-set figure is EXTRAPOLATED and labelled so.
No CUDA in this process: bench.py runs it as a subprocess. Prints one JSON line.
"""
import argparse
import json
import os
import struct
import sys
import time
from multiprocessing import Pool
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF = REPO / "baseline" / "_ref"


def rank(args):                       # evaluation/ranker.py:14-16
    index, qid, query_terms = args
    return qid, index.score(query_terms=query_terms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--index-dir", required=True)
    ap.add_argument("--queries", type=int, default=6980)
    ap.add_argument("--vocab", type=int, default=30522)
    ap.add_argument("--seconds", type=float, default=12.0, help="time budget per leg")
    ap.add_argument("--assumed-postings-per-s-per-core", type=float, default=0.6e6)
    args = ap.parse_args()
    if not (REF / "src").is_dir():
        print(json.dumps({"unavailable": "baseline/_ref/src is missing (run __graft_entry__.build() where /root/reference exists)"}))
        return
    os.chdir(REF)                     # the reference's Logger writes logs/ relative to its own tree
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(REPO))
    from src.deep_impact.inverted_index.inverted_index import InvertedIndex   # the unmodified reference class
    from improving_learned_index_b200 import synthetic

    index = InvertedIndex(index_path=args.index_dir)
    idx = np.fromfile(Path(args.index_dir) / "inverted_index.idx", dtype=np.uint64).reshape(-1, 2)
    df = ((idx[:, 1] - idx[:, 0]) // 5).astype(np.int64)
    queries = synthetic.make_queries(args.queries, vocab_size=args.vocab, seed=7)
    postings = np.array([int(df[[t for t in q if t < len(df)]].sum()) for q in queries])
    cores = max(1, len(os.sched_getaffinity(0)))
    order = np.random.default_rng(123).permutation(len(queries))
    rate = args.assumed_postings_per_s_per_core

    def pick(budget_postings, per_query_cap, max_n):
        chosen, total = [], 0
        for qi in order:
            if 0 < postings[qi] <= per_query_cap and total + postings[qi] <= budget_postings:
                chosen.append(int(qi))
                total += int(postings[qi])
                if len(chosen) == max_n:
                    break
        return chosen

    def terms_of(qi):
        return {synthetic.term_name(t) for t in queries[qi]}      # a Set[str], as process_query returns

    out = {"kind": "reference", "cores": cores, "what": "unmodified InvertedIndex.score from baseline/_ref (inverted_index.py:55-62)",
           "index_files": "vocab.txt / inverted_index.idx / inverted_index.dat written by this repo's builder (byte format of create.py:27-51)",
           "mean_postings_per_query": float(postings.mean())}
    # (i) single process
    s1 = pick(rate * args.seconds, rate * args.seconds / 2, 8)
    t0 = time.perf_counter()
    lens = [len(index.score(terms_of(qi), 1000)) for qi in s1]
    dt = time.perf_counter() - t0
    p1 = int(postings[s1].sum())
    out["single_process"] = {"queries": len(s1), "postings": p1, "seconds": round(dt, 2),
                             "postings_per_s": round(p1 / dt), "queries_per_s_on_sample": round(len(s1) / dt, 4),
                             "extrapolated_queries_per_s_full_set": round(p1 / dt / postings.mean(), 5),
                             "results_returned": lens}
    # (ii) Pool(cores), ranker.py:44-46
    sp = pick(rate * args.seconds * cores, rate * args.seconds / 2, 4 * cores)
    tasks = [(index, qi, terms_of(qi)) for qi in sp]
    t0 = time.perf_counter()
    with Pool(cores) as p:
        done = sum(1 for _ in p.imap_unordered(rank, tasks))
    dt = time.perf_counter() - t0
    pp = int(postings[sp].sum())
    out["pool"] = {"workers": cores, "queries": done, "postings": pp, "seconds": round(dt, 2),
                   "postings_per_s": round(pp / dt), "queries_per_s_on_sample": round(done / dt, 4),
                   "extrapolated_queries_per_s_full_set": round(pp / dt / postings.mean(), 5)}
    out["value"] = out["pool"]["extrapolated_queries_per_s_full_set"]
    out["unit"] = "queries/s (EXTRAPOLATED from postings/s on the sample to the mean query of the full set)"
    out["sample"] = (f"{len(s1)} + {len(sp)} random queries (seed 123) of the {len(queries)} whose posting count fits "
                     f"a {args.seconds:.0f} s budget per leg; full 8.8M-doc index files")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
