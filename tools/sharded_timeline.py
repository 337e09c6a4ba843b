#!/usr/bin/env python
"""Where a sharded search step spends its time (run under torchrun, one rank per GPU):
CUDA-event timing of local search / all-gather / merge / proof read-back / round 2, rank 0 prints the medians.
usage: python -m torch.distributed.run --nproc-per-node N tools/sharded_timeline.py [--docs D] [--queries Q] [--top-k K]"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402  (synthetic shard builder)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=8_841_823)
    ap.add_argument("--queries", type=int, default=6980)
    ap.add_argument("--top-k", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from improving_learned_index_b200 import _native, engine, synthetic
    from improving_learned_index_b200.sharded import ShardedSearcher, shard_range, shard_k
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    _native.set_device(local)
    dev = torch.device(f"cuda:{local}")
    L = _native.lib()
    st = torch.cuda.current_stream().cuda_stream
    V, N, k = 30522, args.docs, args.top_k
    lo, hi = shard_range(N, world, rank)

    def quantize_fn(x):
        out = torch.empty(x.numel(), dtype=torch.int32, device=dev)
        _native.check(L.di_quantize_f64_dev(x.data_ptr(), x.numel(), bench.IMPACT_CLIP, out.data_ptr(), st))
        return out
    terms, imps, offs = bench.build_shard_arrays(lo, hi, N, V, 120, torch, dev, quantize_fn)
    P = terms.numel()
    toff = torch.empty(V + 1, dtype=torch.int64, device=dev)
    docids = torch.empty(P, dtype=torch.int32, device=dev)
    vals = torch.empty(P, dtype=torch.uint8, device=dev)
    _native.check(L.di_invert_dev(terms.data_ptr(), imps.data_ptr(), offs.data_ptr(), hi - lo, V, P, toff.data_ptr(),
                                  docids.data_ptr(), vals.data_ptr(), None, st))
    docids += lo
    index = engine.DeviceIndex.from_csr_device(toff, docids, vals, V, P, doc_lo=lo, doc_hi=hi)
    queries = synthetic.make_queries(args.queries, vocab_size=V, seed=7)
    flat, qoffs = engine.flatten_queries(queries)
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).to(dev)
    d_offs = torch.from_numpy(qoffs.astype(np.int64)).to(dev)
    Q, max_len = len(queries), max(len(q) for q in queries)
    s = ShardedSearcher.for_device_index(index, dev)
    k_in = shard_k(k, world)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    rows = []
    for rep in range(args.reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        keys = s._buf("t_keys", (Q, k_in), torch.int64)
        counts = s._buf("t_counts", (Q,), torch.int32)
        g_keys = s._buf("t_gk", (world, Q, k_in), torch.int64)
        g_counts = s._buf("t_gc", (world, Q), torch.int32)
        out_keys = s._buf("t_ok", (Q, k), torch.int64)
        out_counts = s._buf("t_oc", (Q,), torch.int32)
        inc = s._buf("t_inc", (Q,), torch.int32)
        e0 = ev()
        s.local_search(d_flat, d_offs, Q, max_len, k_in, keys, counts)
        e1 = ev()
        dist.all_gather_into_tensor(g_keys.view(world * Q, k_in), keys)
        dist.all_gather_into_tensor(g_counts.view(world * Q), counts)
        e2 = ev()
        s.merge(g_keys, g_counts, world, Q, k_in, k, out_keys, out_counts, inc)
        e3 = ev()
        n_redo = int(torch.nonzero(inc).numel())
        e4 = ev()
        e0b = ev()
        s.search_tensors(d_flat, d_offs, Q, max_len, k)      # the real two-round call, for the total
        e5 = ev()
        torch.cuda.synchronize()
        if rep >= 2:
            rows.append([e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4),
                         e0b.elapsed_time(e5), n_redo])
    r = np.median(np.array(rows), axis=0)
    t = torch.tensor(r[:5], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} docs={N} Q={Q} k={k} k_in={k_in}: search {t[0]:.3f} ms | all-gather {t[1]:.3f} | merge+proof {t[2]:.3f} | "
              f"flag read-back {t[3]:.3f} | whole two-round call {t[4]:.3f} | redo queries {int(r[5])}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
