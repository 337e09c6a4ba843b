#!/usr/bin/env python
"""How much of a SHARD's search time is threshold warm-up? One GPU holds the first 1/8 of the 8.8 M-document collection (what a
rank of an 8-GPU run holds); the 6,980-query search is timed (a) as the sharded path runs it (per-shard seeds), (b) started from
each query's own final k-th key (the best threshold the shard could ever have), (c) from the k-th key of the GLOBAL top-k
restricted to... not available on one GPU, so (b) is the bound reported. usage: python tools/theta_experiment.py [--shards 8]"""
import argparse
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", type=int, default=8)
    args = ap.parse_args()
    import torch
    from improving_learned_index_b200 import _native, engine, synthetic
    from improving_learned_index_b200.sharded import shard_range, shard_k
    dev = torch.device("cuda:0")
    L = _native.lib()
    st = torch.cuda.current_stream().cuda_stream
    N, V, k = 8_841_823, 30522, 1000
    lo, hi = shard_range(N, args.shards, 0)

    def quantize_fn(x):
        out = torch.empty(x.numel(), dtype=torch.int32, device=dev)
        _native.check(L.di_quantize_f64_dev(x.data_ptr(), x.numel(), bench.IMPACT_CLIP, out.data_ptr(), st))
        return out
    terms, imps, offs = bench.build_shard_arrays(lo, hi, N, V, 208, torch, dev, quantize_fn, 120)
    torch.cuda.synchronize()
    index = engine.DeviceIndex.from_docmajor_device(terms, imps, offs, hi - lo, V, terms.numel(), doc_lo=lo)
    index.set_sorted_prefix(shard_k(k, args.shards))
    queries = synthetic.make_queries(6980, vocab_size=V, seed=7)
    flat, qoffs = engine.flatten_queries(queries)
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).to(dev)
    d_offs = torch.from_numpy(qoffs.astype(np.int64)).to(dev)
    Q, max_len = len(queries), max(len(q) for q in queries)
    keys = torch.zeros((Q, k), dtype=torch.int64, device=dev)
    counts = torch.zeros(Q, dtype=torch.int32, device=dev)

    def timed(theta):
        ms = []
        for _ in range(6):
            index.search_device(d_flat, d_offs, Q, max_len, k, keys, counts, st, d_theta_init=theta)
            torch.cuda.synchronize()
            t = index.timings()
            ms.append((t["score_ms"], t["finalize_ms"]))
        return np.median(np.array(ms[2:]), axis=0)
    a = timed(None)
    index.set_sorted_prefix(0)
    index.search_device(d_flat, d_offs, Q, max_len, k, keys, counts, st)
    torch.cuda.synchronize()
    kth = torch.where(counts == k, keys[:, k - 1], torch.zeros_like(keys[:, 0])).contiguous()
    index.set_sorted_prefix(shard_k(k, args.shards))
    b = timed(kth)
    # (c) GLOBAL single-term seeds: the k-th highest impact of a term over the WHOLE collection (what all-reducing the shards'
    # seed tables at build time would give every shard), from the full term-major inversion
    del terms, imps, offs
    t_all, v_all, o_all = bench.build_shard_arrays(0, N, N, V, 208, torch, dev, quantize_fn, 120)
    P = t_all.numel()
    toff = torch.empty(V + 1, dtype=torch.int64, device=dev)
    docs = torch.empty(P, dtype=torch.int32, device=dev)
    vals = torch.empty(P, dtype=torch.uint8, device=dev)
    _native.check(L.di_invert_dev(t_all.data_ptr(), v_all.data_ptr(), o_all.data_ptr(), N, V, P, toff.data_ptr(), docs.data_ptr(),
                                  vals.data_ptr(), None, st))
    torch.cuda.synchronize()
    del t_all, v_all, o_all, docs
    df = toff[1:] - toff[:-1]
    kth_impact = torch.where(df >= k, vals[(toff[:-1] + k - 1).clamp(max=P - 1)].to(torch.int64), torch.zeros_like(df))
    seeds = torch.zeros(Q, dtype=torch.int64, device=dev)
    for i, q in enumerate(queries):
        seeds[i] = int(kth_impact[torch.tensor(q, device=dev)].max()) << 32
    c = timed(seeds.contiguous())
    print(f"  with GLOBAL single-term seeds (k-th highest impact of a term over all {N} documents): score kernel {c[0]:.3f} ms, finalize {c[1]:.3f}")
    print(f"shard 0 of {args.shards} ({hi - lo} docs, {index.info()['n_tiles']} tiles): score kernel {a[0]:.3f} ms with its own seeds, "
          f"{b[0]:.3f} ms when every query starts from its final k-th key (finalize {a[1]:.3f} / {b[1]:.3f})")


if __name__ == "__main__":
    main()
