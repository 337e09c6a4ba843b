#!/usr/bin/env python
"""Where a sharded search step spends its time (run under torchrun, one rank per GPU): CUDA-event timing of
local search (score kernel + per-shard select/sort) / cross-GPU barrier / fused pull-merge / device-to-host copy of
this rank's slice, next to the NCCL all-gather form of the same step. Rank 0 prints the max-over-ranks medians.
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
    terms, imps, offs = bench.build_shard_arrays(lo, hi, N, V, 208, torch, dev, quantize_fn, 120)
    torch.cuda.synchronize()
    index = engine.DeviceIndex.from_docmajor_device(terms, imps, offs, hi - lo, V, terms.numel(), doc_lo=lo)
    del terms, imps, offs
    queries = synthetic.make_queries(args.queries, vocab_size=V, seed=7)
    flat, qoffs = engine.flatten_queries(queries)
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).to(dev)
    d_offs = torch.from_numpy(qoffs.astype(np.int64)).to(dev)
    Q, max_len = len(queries), max(len(q) for q in queries)
    s = ShardedSearcher.for_device_index(index, dev)
    k_in = shard_k(k, world)
    q_lo, q_hi = shard_range(Q, world, rank)
    n_own = q_hi - q_lo
    h_keys = torch.empty((max(n_own, 1), k), dtype=torch.int64).pin_memory()
    h_counts = torch.empty(max(n_own, 1), dtype=torch.int32).pin_memory()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    s.search_partitioned(d_flat, d_offs, Q, max_len, k)             # sets up the peer buffers (collective)
    ex = s._peer
    own_keys = torch.zeros((max(n_own, 1), k), dtype=torch.int64, device=dev)
    own_counts = torch.zeros(max(n_own, 1), dtype=torch.int32, device=dev)
    second = torch.zeros(1, dtype=torch.int32, device=dev)
    rows = []
    for rep in range(args.reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        # ---- fused form, stage by stage
        r_rows, r_counts, row_table, cnt_table = ex.next_set()
        e0 = ev()
        s.local_search(d_flat, d_offs, Q, max_len, k, r_rows, r_counts)
        t = index.timings()
        e1 = ev()
        ex.barrier(st)
        e2 = ev()
        second.zero_()
        engine.merge_pull_device(row_table, cnt_table, world, q_lo, n_own, k, k_in, k, own_keys, own_counts, st, d_n_second_pass=second)
        e3 = ev()
        h_keys[:n_own].copy_(own_keys[:n_own], non_blocking=True)
        h_counts[:n_own].copy_(own_counts[:n_own], non_blocking=True)
        e4 = ev()
        torch.cuda.synchronize()
        # ---- the NCCL all-gather form of the same step, as one call
        os.environ["DI_B200_NO_PEER"] = "1"
        dist.barrier()
        torch.cuda.synchronize()
        e5 = ev()
        s.search_tensors(d_flat, d_offs, Q, max_len, k)
        e6 = ev()
        torch.cuda.synchronize()
        del os.environ["DI_B200_NO_PEER"]
        if rep >= 2:
            rows.append([e0.elapsed_time(e1), t["score_ms"], t["finalize_ms"], e1.elapsed_time(e2), e2.elapsed_time(e3),
                         e3.elapsed_time(e4), e0.elapsed_time(e4), e5.elapsed_time(e6), float(second.item())])
    r = np.median(np.array(rows), axis=0)
    t = torch.tensor(r, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world={world} docs={N} Q={Q} k={k} k_in={k_in} (max over ranks of per-rank medians, ms)\n"
              f"  fused step        : local search {t[0]:.3f} (score kernel {t[1]:.3f} + per-shard select/sort {t[2]:.3f}) | "
              f"stream barrier incl. waiting for the slowest shard {t[3]:.3f} | pull-merge of {n_own} own queries {t[4]:.3f} "
              f"(second pass for {int(t[8])}) | D2H of the slice {t[5]:.3f} | total {t[6]:.3f}\n"
              f"  NCCL all-gather form of the same step (search + 2 all-gathers of k_in columns + K5 on every rank + full rows of the "
              f"unproven queries): {t[7]:.3f}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
