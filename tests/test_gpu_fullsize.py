"""Full-size checks (BASELINE.json configs[1]/[2]/[4]: 8 841 823 documents, ~8 x 10^8 postings) through the C ABI.
Exact comparison with the oracle on a query sample, plus size-independent properties on all 6 980 queries:
sorted unique keys, idempotence, and shard-and-merge == single index."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from improving_learned_index_b200 import _native, engine, synthetic as syn   # noqa: E402
from oracle import oracle                                                    # noqa: E402

pytestmark = pytest.mark.gpu
N_DOCS, VOCAB, DRAWS, N_QUERIES, K = 8_841_823, 30522, 120, 6980, 1000


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
    terms, imps, offs = bench.build_shard_arrays(0, N_DOCS, N_DOCS, VOCAB, DRAWS, torch, dev, quantize_fn)
    P = terms.numel()
    toff = torch.empty(VOCAB + 1, dtype=torch.int64, device=dev)
    docids = torch.empty(P, dtype=torch.int32, device=dev)
    vals = torch.empty(P, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    _native.check(L.di_invert_dev(terms.data_ptr(), imps.data_ptr(), offs.data_ptr(), N_DOCS, VOCAB, P,
                                  toff.data_ptr(), docids.data_ptr(), vals.data_ptr(), st))
    torch.cuda.synchronize()
    del terms, imps, offs
    index = engine.DeviceIndex.from_csr_device(toff, docids, vals, VOCAB, P)
    queries = syn.make_queries(N_QUERIES, vocab_size=VOCAB, seed=7)
    yield dict(torch=torch, dev=dev, toff=toff, docids=docids, vals=vals, P=P, index=index, queries=queries)
    index.close()


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


def test_sample_matches_oracle_exactly(full):
    h_toff = full["toff"].cpu().numpy().astype(np.uint64)
    h_docs = full["docids"].cpu().numpy().view(np.uint32)
    h_vals = full["vals"].cpu().numpy()
    rng = np.random.default_rng(99)
    pick = sorted(rng.choice(N_QUERIES, size=24, replace=False).tolist())
    sample = [full["queries"][i] for i in pick] + [[], [VOCAB + 3], full["queries"][0] * 2]
    for k in (10, K):
        want = oracle.score_topk_csr(h_toff, h_docs, h_vals, N_DOCS, sample, k)
        got = full["index"].search(sample, k)
        assert np.array_equal(got[2], want[2])
        for i in range(len(sample)):
            n = int(want[2][i])
            assert np.array_equal(got[0][i, :n], want[0][i, :n]) and np.array_equal(got[1][i, :n], want[1][i, :n]), (k, i)


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
