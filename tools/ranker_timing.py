#!/usr/bin/env python
"""Where the wall clock of the drop-in Ranker.run goes (GPU box): prepare / search / write per batch size, on a synthetic
index of --docs documents. Usage: python tools/ranker_timing.py [--docs 2000000]"""
import argparse
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=8_841_823)
    args = ap.parse_args()
    import torch
    import bench
    from improving_learned_index_b200 import _native, engine, synthetic
    from improving_learned_index_b200.utils.datasets import RunFile
    dev = torch.device("cuda:0")
    L = _native.lib()
    st = torch.cuda.current_stream().cuda_stream

    def quantize_fn(x):
        out = torch.empty(x.numel(), dtype=torch.int32, device=dev)
        _native.check(L.di_quantize_f64_dev(x.data_ptr(), x.numel(), bench.IMPACT_CLIP, out.data_ptr(), st))
        return out
    V = 30522
    terms, imps, offs = bench.build_shard_arrays(0, args.docs, args.docs, V, 208, torch, dev, quantize_fn, 120)
    torch.cuda.synchronize()
    index = engine.DeviceIndex.from_docmajor_device(terms, imps, offs, args.docs, V, terms.numel())
    del terms, imps, offs
    queries = synthetic.make_queries(6980, vocab_size=V, seed=7)
    names = [[synthetic.term_name(t) for t in q] for q in queries]
    vocab = {synthetic.term_name(t): t for t in range(V)}
    tmp = Path(tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    os.environ["DI_B200_IO_TRACE"] = "1"
    for B in (1745, 3490, 6980):
        t_prep = t_search = t_write = 0.0
        run = tmp / f"run{B}.tsv"
        for rep in range(2):
            if run.exists():
                run.unlink()
            t_prep = t_search = t_write = 0.0
            for lo in range(0, len(queries), B):
                t0 = time.perf_counter()
                ids = [[vocab.get(t, -1) for t in set(q)] for q in names[lo:lo + B]]
                t1 = time.perf_counter()
                d, s, c = index.search(ids, 1000, pinned=True)
                t2 = time.perf_counter()
                RunFile(run).write_batch([str(i) for i in range(lo, min(lo + B, len(queries)))], d, s, c)
                t3 = time.perf_counter()
                t_prep += t1 - t0
                t_search += t2 - t1
                t_write += t3 - t2
        print(f"batch {B}: prepare {1e3 * t_prep:.1f} ms, search {1e3 * t_search:.1f} ms, write {1e3 * t_write:.1f} ms "
              f"({run.stat().st_size / 1e6:.0f} MB)", flush=True)


if __name__ == "__main__":
    main()
