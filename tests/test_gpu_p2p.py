"""K5 over peer memory (di_merge_rows_p2p_dev), EXPERIMENTAL in round 1: opt-in with DI_B200_P2P=1.

The kernel only dereferences a table of per-shard row pointers, so one GPU is enough to check it: three shards
searched on the same device, pointers to their own result rows, merged without any gather and compared with one
index over all documents — first pass with short rows + proof flags, second pass over the flagged queries only."""
import os

import numpy as np
import pytest

from improving_learned_index_b200 import engine, synthetic as syn
from helpers import assert_same_results, quantized_csr

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DI_B200_P2P") != "1", reason="experimental path: set DI_B200_P2P=1")]


def test_p2p_merge_equals_single_index():
    torch = pytest.importorskip("torch")
    x = quantized_csr(6000, 900, 60, 41)
    # term 900 lives only in documents 0..599 (first shard): its top-k cannot be proven with short rows
    toff = np.concatenate([x["toff"], [x["toff"][-1] + 600]]).astype(np.uint64)
    docs = np.concatenate([x["docs"], np.arange(600, dtype=np.uint32)])
    vals = np.concatenate([x["vals"], (255 - np.arange(600) % 100).astype(np.uint8)])
    full = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=1024)
    queries = syn.make_queries(30, vocab_size=900, seed=6)
    queries[7] = [900]
    queries[3] = []
    k, k_in = 300, 120
    want = full.search(queries, k)
    bounds = [(0, 2000), (2000, 4000), (4000, 6000)]
    flat, offs = engine.flatten_queries(queries)
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    Q, G = len(queries), len(bounds)
    rows = [torch.zeros((Q, k), dtype=torch.int64, device=dev) for _ in bounds]
    counts = [torch.zeros(Q, dtype=torch.int32, device=dev) for _ in bounds]
    for (lo, hi), r, c in zip(bounds, rows, counts):
        shard = engine.DeviceIndex.from_csr(toff, docs, vals, doc_lo=lo, doc_hi=hi, tile_docs=1024)
        shard.search_device(d_flat, d_offs, Q, max(len(q) for q in queries), k, r, c, st)
    row_ptrs = torch.tensor([r.data_ptr() for r in rows], dtype=torch.int64, device=dev)
    cnt_ptrs = torch.tensor([c.data_ptr() for c in counts], dtype=torch.int64, device=dev)
    out_keys = torch.zeros((Q, k), dtype=torch.int64, device=dev)
    out_counts = torch.zeros(Q, dtype=torch.int32, device=dev)
    flags = torch.full((Q,), 7, dtype=torch.int32, device=dev)
    engine.merge_rows_p2p_device(row_ptrs, cnt_ptrs, G, Q, k, k_in, k, out_keys, out_counts, st, d_incomplete=flags)
    torch.cuda.synchronize()
    assert set(flags.tolist()) <= {0, 1} and flags[7] == 1
    redo = torch.nonzero(flags).flatten()
    k2 = torch.zeros((int(redo.numel()), k), dtype=torch.int64, device=dev)
    c2 = torch.zeros(int(redo.numel()), dtype=torch.int32, device=dev)
    engine.merge_rows_p2p_device(row_ptrs, cnt_ptrs, G, int(redo.numel()), k, k, k, k2, c2, st, d_query_ids=redo.to(torch.int32))
    out_keys[redo] = k2
    out_counts[redo] = c2
    torch.cuda.synchronize()
    keys_np = out_keys.cpu().numpy().view(np.uint64)
    got = ((~(keys_np & np.uint64(0xFFFFFFFF)).astype(np.uint32)), (keys_np >> np.uint64(32)).astype(np.int32),
           out_counts.cpu().numpy().view(np.uint32))
    assert_same_results(got, want, "p2p merge")
