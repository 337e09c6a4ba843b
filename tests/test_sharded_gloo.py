"""world_size-2 test of the sharded search plumbing on CPU (gloo): shard ranges, the [G, Q, k] layout
of the all-gather, count handling and the global order. The two compute steps are test doubles built on
the oracle (the CUDA steps are covered by test_gpu_parity.py::test_shards_and_merge_equal_single_index)."""
import os
import socket

import numpy as np
import pytest

from improving_learned_index_b200 import synthetic as syn
from improving_learned_index_b200.sharded import ShardedSearcher, pack_keys, shard_k, shard_range, unpack_keys
from oracle import oracle
from helpers import quantized_csr

torch = pytest.importorskip("torch")


def test_shard_range_covers_everything():
    for n, g in ((10, 3), (8_841_823, 8), (5, 8), (0, 2), (7, 1)):
        spans = [shard_range(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(hi - lo <= -(-n // g) for lo, hi in spans) if n else True


def test_key_packing_orders_by_score_then_docid():
    scores = np.array([5, 5, 7, 1], dtype=np.int32)
    docs = np.array([9, 2, 100, 0], dtype=np.uint32)
    keys = pack_keys(scores, docs)
    order = np.argsort(keys)[::-1]
    assert docs[order].tolist() == [100, 2, 9, 0]
    s, d = unpack_keys(keys)
    assert np.array_equal(s, scores) and np.array_equal(d, docs)


def _worker(rank, world, port, out_dict):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = quantized_csr(4000, 600, 50, 77)
    # term 600: only documents 0..699 (all inside shard 0) carry it, with their best impacts -> a query on it has
    # its whole top-k in one shard, which round 1 (k_in < k keys per shard) cannot prove complete
    extra_docs = np.arange(700, dtype=np.uint32)
    extra_vals = (255 - (extra_docs % 200)).astype(np.uint8)
    toff_full = np.concatenate([x["toff"], [x["toff"][-1] + extra_docs.size]]).astype(np.uint64)
    docs_full = np.concatenate([x["docs"], extra_docs])
    vals_full = np.concatenate([x["vals"], extra_vals])
    n_terms = 601
    lo, hi = shard_range(x["n_docs"], world, rank)
    term_of = np.repeat(np.arange(n_terms), np.diff(toff_full.astype(np.int64)))
    sel = (docs_full >= lo) & (docs_full < hi)
    toff = np.zeros(n_terms + 1, dtype=np.uint64)
    toff[1:] = np.cumsum(np.bincount(term_of[sel], minlength=n_terms))
    docs, vals = docs_full[sel], vals_full[sel]
    queries = syn.make_queries(25, vocab_size=600, seed=4)
    queries[3] = []
    queries[5] = [600]
    queries[6] = [600, queries[6][0]]
    calls = []

    def local_search(qt, qo, n_q, max_len, k, out_keys, out_counts, theta_init=None):
        offs = qo.numpy().astype(np.int64)
        flat = qt.numpy().view(np.uint32)
        qs = [flat[offs[i]:offs[i + 1]].tolist() for i in range(n_q)]
        calls.append((n_q, k))
        d, s, c, _ = oracle.score_topk_csr(toff, docs, vals, x["n_docs"], qs, k)
        keys = np.zeros((n_q, k), dtype=np.uint64)
        for i in range(n_q):
            keys[i, :c[i]] = pack_keys(s[i, :c[i]], d[i, :c[i]])
        if theta_init is not None:        # like the CUDA path: nothing below the proven bound is returned
            th = theta_init.numpy().view(np.uint64)
            for i in range(n_q):
                keep = keys[i, :c[i]] >= th[i]
                c[i] = int(keep.sum())
                keys[i, c[i]:] = 0
        out_keys.copy_(torch.from_numpy(keys.view(np.int64)))
        out_counts.copy_(torch.from_numpy(c.astype(np.int32)))

    def merge(g_keys, g_counts, n_shards, n_q, k_in, k, out_keys, out_counts, incomplete):
        gk, gc = g_keys.numpy().view(np.uint64), g_counts.numpy()
        assert gk.shape == (n_shards, n_q, k_in) and gc.shape == (n_shards, n_q)
        for q in range(n_q):
            allk = np.concatenate([gk[s, q, :gc[s, q]] for s in range(n_shards)])
            top = np.sort(allk)[::-1][:k]
            out_keys[q, :top.size] = torch.from_numpy(top.copy().view(np.int64))
            out_counts[q] = top.size
            kth = top[-1] if top.size == k else np.uint64(0)
            incomplete[q] = int(any(gc[s, q] == k_in and gk[s, q, k_in - 1] > kth for s in range(n_shards)))

    searcher = ShardedSearcher(local_search, merge, torch.device("cpu"), rows_per_shard=lambda k: 5 * k // 8 + 64)
    ok = True
    for k in (50, 400):
        d, s, c = searcher.search(queries, k)
        want = oracle.score_topk_csr(toff_full, docs_full, vals_full, x["n_docs"], queries, k)
        ok = ok and np.array_equal(c, want[2])
        for i in range(len(queries)):
            ok = ok and np.array_equal(d[i, :c[i]], want[0][i, :c[i]]) and np.array_equal(s[i, :c[i]], want[1][i, :c[i]])
    # k = 50: rows are already full-size (one gather); k = 400: 314 columns first, then the full rows of the two
    # clustered queries; every shard searches exactly once per call
    ok = ok and shard_k(400, 2) == 400 and shard_k(1000, 8) == 221 and shard_k(10, 8) == 10 and searcher.round2_queries == 2
    ok = ok and calls == [(25, 50), (25, 400)]
    out_dict[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_world2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    oracle.lib()                      # build the oracle before forking workers
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
