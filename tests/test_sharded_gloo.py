"""world_size-2 test of the sharded search plumbing on CPU (gloo): shard ranges, the [G, Q, k] layout
of the all-gather, count handling and the global order. The two compute steps are test doubles built on
the oracle (the CUDA steps are covered by test_gpu_parity.py::test_shards_and_merge_equal_single_index)."""
import os
import socket

import numpy as np
import pytest

from improving_learned_index_b200 import synthetic as syn
from improving_learned_index_b200.sharded import ShardedSearcher, pack_keys, shard_range, unpack_keys
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
    lo, hi = shard_range(x["n_docs"], world, rank)
    term_of = np.repeat(np.arange(600), np.diff(x["toff"].astype(np.int64)))
    sel = (x["docs"] >= lo) & (x["docs"] < hi)
    toff = np.zeros(601, dtype=np.uint64)
    toff[1:] = np.cumsum(np.bincount(term_of[sel], minlength=600))
    docs, vals = x["docs"][sel], x["vals"][sel]
    queries = syn.make_queries(25, vocab_size=600, seed=4)
    queries[3] = []

    def local_search(qt, qo, n_q, max_len, k, out_keys, out_counts):
        d, s, c, _ = oracle.score_topk_csr(toff, docs, vals, x["n_docs"], queries, k)
        keys = np.zeros((n_q, k), dtype=np.uint64)
        for i in range(n_q):
            keys[i, :c[i]] = pack_keys(s[i, :c[i]], d[i, :c[i]])
        out_keys.copy_(torch.from_numpy(keys.view(np.int64)))
        out_counts.copy_(torch.from_numpy(c.astype(np.int32)))

    def merge(g_keys, g_counts, n_shards, n_q, k, out_keys, out_counts):
        gk, gc = g_keys.numpy().view(np.uint64), g_counts.numpy()
        assert gk.shape == (n_shards, n_q, k) and gc.shape == (n_shards, n_q)
        for q in range(n_q):
            allk = np.concatenate([gk[s, q, :gc[s, q]] for s in range(n_shards)])
            top = np.sort(allk)[::-1][:k]
            out_keys[q, :top.size] = torch.from_numpy(top.copy().view(np.int64))
            out_counts[q] = top.size

    searcher = ShardedSearcher(local_search, merge, torch.device("cpu"))
    d, s, c = searcher.search(queries, 50)
    want = oracle.score_topk_csr(x["toff"], x["docs"], x["vals"], x["n_docs"], queries, 50)
    ok = np.array_equal(c, want[2])
    for i in range(len(queries)):
        ok = ok and np.array_equal(d[i, :c[i]], want[0][i, :c[i]]) and np.array_equal(s[i, :c[i]], want[1][i, :c[i]])
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
