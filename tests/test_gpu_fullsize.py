"""Full-size checks (BASELINE.json configs[1]..[4]: 8 841 823 documents x 120 distinct terms, ~1.05 x 10^9 postings)
through the C ABI: the whole K2 inversion against the oracle's (configs[4]), exact comparison with the oracle on
query samples drawn from full batches (6 980 queries top-1000 = configs[1]; 4 096 queries top-100 = configs[3]), plus
size-independent properties on all 6 980 queries: sorted unique keys, idempotence, shard-and-merge == single index."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from improving_learned_index_b200 import _native, engine, synthetic as syn   # noqa: E402
from oracle import oracle                                                    # noqa: E402

pytestmark = pytest.mark.gpu
N_DOCS, VOCAB, DRAWS, UNIQUE, N_QUERIES, K = 8_841_823, 30522, 208, 120, 6980, 1000


@pytest.fixture(scope="module")
def full():
    torch = pytest.importorskip("torch")
    import bench
    dev = torch.device("cuda:0")
    L = _native.lib()
    st = torch.cuda.current_stream().cuda_stream

    def quantize_fn(x):
        out = torch.empty(x.numel(), dtype=torch.int32, device=dev)
        _native.check(L.di_quantize_f64_dev(x.data_ptr(), x.numel(), bench.IMPACT_CLIP, out.data_ptr(), st))
        return out
    terms, imps, offs = bench.build_shard_arrays(0, N_DOCS, N_DOCS, VOCAB, DRAWS, torch, dev, quantize_fn, UNIQUE)
    P = terms.numel()
    toff = torch.empty(VOCAB + 1, dtype=torch.int64, device=dev)
    docids = torch.empty(P, dtype=torch.int32, device=dev)
    vals = torch.empty(P, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    _native.check(L.di_invert_dev(terms.data_ptr(), imps.data_ptr(), offs.data_ptr(), N_DOCS, VOCAB, P,
                                  toff.data_ptr(), docids.data_ptr(), vals.data_ptr(), None, st))
    torch.cuda.synchronize()
    # the oracle's own inversion of the same doc-major arrays (create.py:31-46 restated, ~20 s single-threaded): the
    # reference for configs[4] and the CSR every scoring comparison below runs the oracle on
    o_toff, o_docs, o_vals = oracle.invert(terms.cpu().numpy().view(np.uint32), imps.cpu().numpy(),
                                           offs.cpu().numpy().astype(np.uint64), VOCAB)
    del terms, imps, offs
    index = engine.DeviceIndex.from_csr_device(toff, docids, vals, VOCAB, P)
    queries = syn.make_queries(N_QUERIES, vocab_size=VOCAB, seed=7)
    yield dict(torch=torch, dev=dev, toff=toff, docids=docids, vals=vals, P=P, index=index, queries=queries,
               o_toff=o_toff, o_docs=o_docs, o_vals=o_vals)
    index.close()


def test_full_size_inversion_is_bit_exact(full):
    """configs[4]: quantize + term->doc inversion of the 8.8 M per-document lists, every posting compared."""
    assert full["P"] > 1_000_000_000                               # ~120 distinct terms per document survive K1
    assert np.array_equal(full["toff"].cpu().numpy().astype(np.uint64), full["o_toff"])
    assert np.array_equal(full["docids"].cpu().numpy().view(np.uint32), full["o_docs"])
    assert np.array_equal(full["vals"].cpu().numpy(), full["o_vals"])


def test_csr_is_well_formed(full):
    torch = full["torch"]
    toff = full["toff"]
    assert int(toff[0]) == 0 and int(toff[-1]) == full["P"] and bool((toff[1:] >= toff[:-1]).all())
    # inside each term's list: impact non-increasing (create.py:41); checked on the 20 most frequent terms
    df = (toff[1:] - toff[:-1])
    for t in torch.topk(df, 20).indices.tolist():
        v = full["vals"][int(toff[t]):int(toff[t + 1])].to(torch.int16)
        d = full["docids"][int(toff[t]):int(toff[t + 1])].to(torch.int64)
        assert bool((v[1:] <= v[:-1]).all())
        same = v[1:] == v[:-1]
        assert bool((d[1:][same] > d[:-1][same]).all())          # ties by ascending docid
    assert int(full["vals"].min()) >= 1                          # quantize.py:45 dropped the zeros


def _check_rows(full, got, queries, pick, k):
    sample = [queries[i] for i in pick]
    want = oracle.score_topk_csr(full["o_toff"], full["o_docs"], full["o_vals"], N_DOCS, sample, k)
    for i, qi in enumerate(pick):
        n = int(want[2][i])
        assert int(got[2][qi]) == n, (k, qi)
        assert np.array_equal(got[0][qi, :n], want[0][i, :n]) and np.array_equal(got[1][qi, :n], want[1][i, :n]), (k, qi)


def test_timed_batch_rows_match_oracle_exactly(full):
    """configs[1]: ALL 6 980 queries in one call (one tile chain per query, 540 hand-offs each — the timed path);
    the rows of 24 random queries plus the edge cases are compared with the oracle."""
    queries = list(full["queries"])
    queries[11], queries[12], queries[13] = [], [VOCAB + 3], queries[0] * 2
    rng = np.random.default_rng(99)
    pick = sorted(set(rng.choice(N_QUERIES, size=24, replace=False).tolist()) | {11, 12, 13})
    for k in (10, K):
        got = full["index"].search(queries, k)
        assert full["index"].timings()["lanes"] == 1
        _check_rows(full, got, queries, pick, k)
    small = [queries[i] for i in pick[:8]]                       # the same queries as a small batch (tile lanes)
    got = full["index"].search(small, K)
    assert full["index"].timings()["lanes"] > 1
    _check_rows(full, got, small, list(range(len(small))), K)


def test_c4_batch_rows_match_oracle_exactly(full):
    """configs[3]: a batch of 4 096 queries of ~6 terms, top-100."""
    queries = syn.make_queries(4096, vocab_size=VOCAB, seed=11)
    got = full["index"].search(queries, 100)
    assert full["index"].timings()["lanes"] == 1
    pick = sorted(np.random.default_rng(5).choice(4096, size=32, replace=False).tolist())
    _check_rows(full, got, queries, pick, 100)
    counts = got[2]
    keys = (got[1].astype(np.uint64) << np.uint64(32)) | (~got[0]).astype(np.uint64)
    for i in range(0, 4096, 5):
        assert np.all(keys[i, 1:counts[i]] < keys[i, :counts[i] - 1])


def test_all_queries_properties_and_shard_merge(full):
    torch, dev, index, queries = full["torch"], full["dev"], full["index"], full["queries"]
    d1, s1, c1 = index.search(queries, K)
    d2, s2, c2 = index.search(queries, K)                        # idempotent, deterministic
    assert np.array_equal(c1, c2) and np.array_equal(d1, d2) and np.array_equal(s1, s2)
    keys = (s1.astype(np.uint64) << np.uint64(32)) | (~d1).astype(np.uint64)
    for i in range(0, N_QUERIES, 7):
        n = int(c1[i])
        assert n == K or n < K
        assert np.all(keys[i, 1:n] < keys[i, :n - 1])            # strictly descending: sorted, no duplicate doc
        assert s1[i, :n].min(initial=1) >= 1 and d1[i, :n].max(initial=0) < N_DOCS
    # upper bound on any score: 255 per query term occurrence
    lens = np.array([len(q) for q in queries])
    assert np.all(s1.max(axis=1) <= 255 * lens)
    # docid-range shards + K5 merge reproduce the single-index result for every query
    bounds = [(0, 3_000_000), (3_000_000, 6_500_000), (6_500_000, N_DOCS)]
    flat, offs = engine.flatten_queries(queries)
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(dev).to(torch.int32)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    g_keys = torch.zeros((3, N_QUERIES, K), dtype=torch.int64, device=dev)
    g_counts = torch.zeros((3, N_QUERIES), dtype=torch.int32, device=dev)
    total = 0
    for s, (lo, hi) in enumerate(bounds):
        shard = engine.DeviceIndex.from_csr_device(full["toff"], full["docids"], full["vals"], VOCAB, full["P"],
                                                   doc_lo=lo, doc_hi=hi)
        total += shard.info()["n_postings"]
        shard.search_device(d_flat, d_offs, N_QUERIES, int(lens.max()), K, g_keys[s], g_counts[s], st)
        torch.cuda.synchronize()
        shard.close()
    assert total == index.info()["n_postings"] == full["P"]
    out_keys = torch.zeros((N_QUERIES, K), dtype=torch.int64, device=dev)
    out_counts = torch.zeros(N_QUERIES, dtype=torch.int32, device=dev)
    engine.merge_topk_device(g_keys, g_counts, 3, N_QUERIES, K, out_keys, out_counts, st)
    torch.cuda.synchronize()
    assert np.array_equal(out_counts.cpu().numpy().view(np.uint32), c1)
    merged = out_keys.cpu().numpy().view(np.uint64)
    for i in range(N_QUERIES):
        assert np.array_equal(merged[i, :c1[i]], keys[i, :c1[i]]), i
