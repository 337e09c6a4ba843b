"""K5 fused with its exchange (di_merge_pull_dev) and the stream barrier (di_peer_barrier_dev).

The merge kernel only dereferences a table of per-shard row pointers, so ONE GPU is enough to check its logic: three
shards searched on the same device, pointers to their own result rows, every "rank" merging its slice of the queries
(pull of short rows, proof, in-kernel second pass with full rows), compared with one index over all documents. The real
thing — rows in another process's GPU memory, mapped through CUDA IPC and read over NVLink — needs two GPUs:
test_two_process_peer_exchange runs wherever at least two are visible (gpurun --gpus 2)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from improving_learned_index_b200 import engine, synthetic as syn
from improving_learned_index_b200.sharded import shard_range
from helpers import assert_same_results, quantized_csr

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _clustered_collection():
    x = quantized_csr(6000, 900, 60, 41)
    # term 900 lives only in documents 0..599 (first shard): its top-k cannot be proven with short rows
    toff = np.concatenate([x["toff"], [x["toff"][-1] + 600]]).astype(np.uint64)
    docs = np.concatenate([x["docs"], np.arange(600, dtype=np.uint32)])
    vals = np.concatenate([x["vals"], (255 - np.arange(600) % 100).astype(np.uint8)])
    queries = syn.make_queries(30, vocab_size=900, seed=6)
    queries[7] = [900]
    queries[3] = []
    return toff, docs, vals, queries


@pytest.mark.parametrize("shared_seeds", [False, True])
@pytest.mark.parametrize("k,k_in", [(300, 120), (300, 300), (1000, 221), (10, 4)])
def test_pull_merge_equals_single_index(k, k_in, shared_seeds):
    torch = pytest.importorskip("torch")
    toff, docs, vals, queries = _clustered_collection()
    full = engine.DeviceIndex.from_csr(toff, docs, vals, tile_docs=1024)
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
    shards = [engine.DeviceIndex.from_csr(toff, docs, vals, doc_lo=lo, doc_hi=hi, tile_docs=1024) for lo, hi in bounds]
    if shared_seeds:     # every shard gets the seed tables of the whole collection (sum of the shards' impact histograms)
        total = torch.zeros((toff.size - 1, 256), dtype=torch.int32, device=dev)
        for shard in shards:
            h = torch.zeros_like(total)
            shard.export_seed_hist(h, st)
            total += h
        torch.cuda.synchronize()
        whole = torch.zeros_like(total)
        full.export_seed_hist(whole, st)
        torch.cuda.synchronize()
        assert torch.equal(total, whole)                                    # shards partition the postings
        for shard in shards:
            shard.import_seed_hist(total, st)
    for shard, r, c in zip(shards, rows, counts):
        shard.set_sorted_prefix(k_in if k_in < k else 1)                    # what a shard owes the merge, no more
        shard.search_device(d_flat, d_offs, Q, max(len(q) for q in queries), k, r, c, st)
    row_ptrs = torch.tensor([r.data_ptr() for r in rows], dtype=torch.int64, device=dev)
    cnt_ptrs = torch.tensor([c.data_ptr() for c in counts], dtype=torch.int64, device=dev)
    out_keys = torch.zeros((Q, k), dtype=torch.int64, device=dev)
    out_counts = torch.zeros(Q, dtype=torch.int32, device=dev)
    second = torch.zeros(1, dtype=torch.int32, device=dev)
    for rank in range(G):                                    # every "rank" merges its own slice of the queries
        q_lo, q_hi = shard_range(Q, G, rank)
        engine.merge_pull_device(row_ptrs, cnt_ptrs, G, q_lo, q_hi - q_lo, k, k_in, k, out_keys[q_lo:q_hi],
                                 out_counts[q_lo:q_hi], st, d_n_second_pass=second)
    torch.cuda.synchronize()
    keys_np = out_keys.cpu().numpy().view(np.uint64)
    got = ((~(keys_np & np.uint64(0xFFFFFFFF)).astype(np.uint32)), (keys_np >> np.uint64(32)).astype(np.int32),
           out_counts.cpu().numpy().view(np.uint32))
    assert_same_results(got, want, "pull merge")
    if k_in < k:
        assert int(second.item()) >= 1                       # query 7 must have needed the full rows
    else:
        assert int(second.item()) == 0


def test_stream_barrier_orders_three_streams():
    """Three "ranks" = three streams of one GPU; the flag arrays are plain device buffers. Run in a subprocess: a
    broken barrier traps the kernel (by design) and would poison this process's CUDA context."""
    code = textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, %r)
        from improving_learned_index_b200 import engine
        dev = torch.device('cuda:0')
        G = 3
        flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(G)]
        table = torch.tensor([f.data_ptr() for f in flags], dtype=torch.int64, device=dev)
        data = [torch.zeros(1 << 20, dtype=torch.int32, device=dev) for _ in range(G)]
        seen = torch.zeros((G, G), dtype=torch.int32, device=dev)
        streams = [torch.cuda.Stream() for _ in range(G)]
        torch.cuda.synchronize()
        for epoch in range(1, 6):
            for r, s in enumerate(streams):
                with torch.cuda.stream(s):
                    data[r].fill_(epoch)                                   # "search" of rank r
                    engine.peer_barrier_device(table, G, r, epoch, s.cuda_stream)
                    for o in range(G):                                     # "merge": reads every rank's data
                        seen[r, o] = data[o][-1]
            torch.cuda.synchronize()
            assert seen.eq(epoch).all().item(), (epoch, seen.tolist())
        print('barrier ok')
    """) % REPO
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "barrier ok" in r.stdout, r.stdout + r.stderr


_WORKER = """
import os, sys
import numpy as np
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, 'tests'))
import torch
import torch.distributed as dist
from improving_learned_index_b200 import _native, engine, synthetic as syn
from improving_learned_index_b200.sharded import ShardedSearcher, shard_range, unpack_keys
from oracle import oracle
from helpers import quantized_csr

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
_native.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
x = quantized_csr(30000, 1500, 60, 91)
lo, hi = shard_range(30000, world, rank)
shard = engine.DeviceIndex.from_csr(x['toff'], x['docs'], x['vals'], doc_lo=lo, doc_hi=hi, tile_docs=1024)
searcher = ShardedSearcher.for_device_index(shard, dev)
assert searcher.peer_exchange_available()
searcher.share_seeds(shard)           # collective: seed tables of the whole collection on every shard
ok = True
for it, (nq, k) in enumerate([(200, 100), (200, 100), (1200, 300), (1200, 300), (333, 1000), (333, 1000), (50, 7)]):   # 1200 > resident CTAs: one lane, prefix-sorted rows
    queries = syn.make_queries(nq, vocab_size=1500, seed=100 + it)
    queries[0] = []
    searcher.rows_per_shard = (lambda kk: max(1, kk // 3)) if it %% 2 == 0 else None     # short rows force second passes
    d, s, c = searcher.search(queries, k)                      # fused exchange + all-gather of the slices
    want = oracle.score_topk_csr(x['toff'], x['docs'], x['vals'], 30000, queries, k)
    ok = ok and np.array_equal(c, want[2])
    for i in range(nq):
        n = int(want[2][i])
        ok = ok and np.array_equal(d[i, :n], want[0][i, :n]) and np.array_equal(s[i, :n], want[1][i, :n])
    # and the slice form, with no collective at all
    flat, offs = engine.flatten_queries(queries)
    d_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).to(dev)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    (q_lo, q_hi), keys, counts = searcher.search_partitioned(d_flat, d_offs, nq, max(len(q) for q in queries), k)
    torch.cuda.synchronize()
    sc, dd = unpack_keys(keys.cpu().numpy().view(np.uint64))
    cc = counts.cpu().numpy()
    for i in range(q_lo, q_hi):
        n = int(want[2][i])
        ok = ok and int(cc[i - q_lo]) == n and np.array_equal(dd[i - q_lo, :n], want[0][i, :n]) and np.array_equal(sc[i - q_lo, :n], want[1][i, :n])
open(os.path.join(%r, 'rank%%d.%%s' %% (rank, 'ok' if ok else 'mismatch')), 'w').close()    # stdout of the ranks interleaves
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_two_process_peer_exchange(tmp_path):
    """One process per GPU, rows in the other process's memory (CUDA IPC), pulled over NVLink by the merge kernel."""
    torch = pytest.importorskip("torch")
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % (REPO, REPO, str(tmp_path)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=900)
    verdicts = sorted(f.name for f in tmp_path.glob("rank*.*"))
    assert r.returncode == 0 and verdicts == [f"rank{i}.ok" for i in range(world)], (verdicts, r.stdout[-3000:] + r.stderr[-3000:])
