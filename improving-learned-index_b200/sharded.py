"""Docid-range sharding across the GPUs of one box (one process per GPU, torch.distributed).

New functionality: the reference searches on one host only (its parallelism is a per-query
process pool, ranker.py:44-46). Documents are split into `world_size` contiguous docid ranges;
rank r holds a full-vocabulary shard of range r with GLOBAL docids. Every rank scores every
query on its shard (no data-path collective: a document's score depends only on its own
postings), then the per-shard best keys — (score << 32 | ~docid), sorted — are all-gathered
(NCCL over NVLink on GPUs; gloo in the CPU tests) and merged by K5 into the global top-k, which
is identical to the single-GPU result because the key order is total.

Two gathers keep it exact AND cheap. Every shard selects its own top-k once, but only the first k_in < k
columns of its (sorted) rows are gathered at first (about 1.25 k / G + 64: with G shards each holds ~k/G of
the global top-k), which cuts the gathered bytes and the merge work by ~G/1.25. The merge proves the result
complete per query: a shard that filled its k_in columns could only hide keys below the last one it sent, so
the merged top-k is exact iff that key is <= the merged k-th (merge_check_kernel). For the queries that fail
the proof (with docid-ordered ties the boundary tie group of a short query sits in the lowest docid range,
so one shard holds most of the top-k) the full rows — already computed — are gathered and merged; nothing
is searched twice.

On GPUs the exchange is ONE fused kernel over peer memory and no collective at all (`search_partitioned`): every
rank's search writes its sorted rows into a CUDA-IPC buffer that all peers have mapped (`PeerExchange`), a one-CTA
flag barrier orders the searches (`di_peer_barrier_dev`), and `di_merge_pull_dev` lets rank r merge ITS slice of the
queries by pulling the first k_in keys of every shard's row over NVLink straight into shared memory, proving the
result, and pulling the full rows of the queries that fail the proof — inside the same kernel, without a host
round trip. Merge work, pulled bytes and the device-to-host copy of the results are all divided by the number of GPUs.
The all-gather form below stays as the transport-agnostic fallback (gloo in the CPU tests, or when IPC is unavailable).

The collective and the two compute steps are injected, so the plumbing (ranges, tensor layout of
the gather, count handling, the two-round protocol) is testable on CPU with world_size 2 while the
CUDA path plugs in DeviceIndex.search_device and engine.merge_topk_device.
"""
from __future__ import annotations

import os
from typing import Callable, Sequence, Tuple

import numpy as np

from . import engine


def shard_range(n_docs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous docid range [lo, hi) of `rank`: ceil(n_docs / world_size) docs per shard."""
    per = -(-n_docs // world_size)
    return min(rank * per, n_docs), min((rank + 1) * per, n_docs)


def shard_k(k: int, world_size: int) -> int:
    """Keys each shard returns in round 1."""
    if world_size <= 2:   # with two shards the rows would shrink by < 1/3: not worth a possible second round
        return k
    return min(k, -(-5 * k // (4 * world_size)) + 64)


def pack_keys(scores: np.ndarray, docids: np.ndarray) -> np.ndarray:
    """(score, docid) -> ranking key; larger key = better, equal scores order by ascending docid."""
    return (scores.astype(np.uint64) << np.uint64(32)) | (~docids.astype(np.uint32)).astype(np.uint64)


def unpack_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    k = keys.astype(np.uint64)
    return (k >> np.uint64(32)).astype(np.int32), ~(k & np.uint64(0xFFFFFFFF)).astype(np.uint32)


class PeerExchange:
    """This rank's peer-visible result buffers and the tables of every rank's: two sets of (rows [q_cap, k] u64,
    counts [q_cap] u32) that alternate between calls — the barrier of call t + 1 is what protects set t from being
    overwritten while a peer still reads it — and one flag array for the stream barrier. Memory comes from the
    library (cudaMalloc + CUDA IPC handle); torch.distributed only carries the 64-byte handles once."""

    def __init__(self, device, rank: int, world: int, q_cap: int, k: int, group=None):
        """Collective. Raises RuntimeError ON EVERY RANK when any rank could not allocate, export or map the buffers
        (no CUDA IPC / peer access on this box): the ranks agree on the outcome before anybody relies on it."""
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _native as N
        self.N, self.torch = N, torch
        self.rank, self.world, self.q_cap, self.k = rank, world, q_cap, k
        self.calls = self.epoch = 0
        self.own, self.opened = [], []
        L = N.lib()
        sizes = [q_cap * k * 8, q_cap * 4, q_cap * k * 8, q_cap * 4, max(world, 64) * 4]     # rows0 counts0 rows1 counts1 flags
        handles, error = [], None
        try:
            for nbytes in sizes:
                ptr, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
                N.check(L.di_shared_alloc(nbytes, ctypes.byref(ptr), h))
                self.own.append(ptr.value)
                handles.append(h.raw)
        except Exception as e:       # keep going: the collectives below must be entered by every rank
            error = e
        everyone = [None] * world
        dist.all_gather_object(everyone, None if error else handles, group=group)
        table = []
        if error is None and all(h is not None for h in everyone):
            try:
                for r in range(world):
                    if r == rank:
                        table.append(list(self.own))
                        continue
                    ptrs = []
                    for h in everyone[r]:
                        ptr = ctypes.c_void_p()
                        N.check(L.di_shared_open(h, ctypes.byref(ptr)))
                        ptrs.append(ptr.value)
                        self.opened.append(ptr.value)
                    table.append(ptrs)
            except Exception as e:
                error = e
        elif error is None:
            error = RuntimeError("a peer rank could not allocate its buffers")
        ok = torch.tensor([0 if error else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also: nobody starts a barrier kernel before every rank has mapped the flags
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"peer-memory exchange unavailable ({error or 'failure on another rank'})")
        as_table = lambda j: torch.tensor([table[r][j] for r in range(world)], dtype=torch.int64, device=device)
        self.sets = [(self.own[0], self.own[1], as_table(0), as_table(1)), (self.own[2], self.own[3], as_table(2), as_table(3))]
        self.flag_table = as_table(4)

    def next_set(self):
        s = self.sets[self.calls % 2]
        self.calls += 1
        return s

    def barrier(self, stream: int):
        self.epoch += 1
        self.N.check(self.N.lib().di_peer_barrier_dev(self.flag_table.data_ptr(), self.world, self.rank, self.epoch, stream))

    def close(self):
        L = self.N.lib()
        self.torch.cuda.synchronize()
        for p in self.opened:
            L.di_shared_close(p)
        for p in self.own:
            L.di_shared_free(p)
        self.opened, self.own = [], []


class ShardedSearcher:
    """search(): local best keys on this rank's shard -> all_gather of the first columns -> merge + proof
    (-> all_gather of the full rows of the unproven queries -> merge).
    Call on every rank with the same queries; every rank gets the global result.

    local_search(d_q_terms, d_q_offsets, n_queries, max_len, k, out_keys, out_counts[, theta_init=]) fills this
    shard's sorted keys [Q, k] (int64 bit patterns) and counts [Q] (int32) — torch tensors on `device`;
    theta_init [Q] (int64 keys), when given, are proven lower bounds of the final k-th keys: the shard may
    leave out anything below them.
    merge(gathered_keys [G,Q,k_in], gathered_counts [G,Q], G, Q, k_in, k, out_keys [Q,k], out_counts [Q],
    incomplete [Q] int32) writes the global top-k and flags the queries whose merge is not proven exact.
    """

    def __init__(self, local_search: Callable, merge: Callable, device, group=None, rows_per_shard=None, set_row_order=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.local_search, self.merge = local_search, merge
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._buffers = {}
        self.rows_per_shard = rows_per_shard   # optional override of shard_k: k -> keys per shard in round 1
        self.round2_queries = 0          # how many queries of the last search needed their full rows gathered
        self.set_row_order = set_row_order   # optional: p -> "only the first p keys of a local row need to be sorted"
        self._peer = None                # PeerExchange of the fused path (CUDA, world > 1), created on first use
        self._peer_ok = None             # None = not tried yet
        self._second = None

    @classmethod
    def for_device_index(cls, index: "engine.DeviceIndex", device, group=None) -> "ShardedSearcher":
        import torch

        def local_search(qt, qo, n_q, max_len, k, out_keys, out_counts, theta_init=None):
            index.search_device(qt, qo, n_q, max_len, k, out_keys, out_counts, torch.cuda.current_stream().cuda_stream,
                                d_theta_init=theta_init)

        def merge(g_keys, g_counts, n_shards, n_q, k_in, k, out_keys, out_counts, incomplete):
            engine.merge_topk_device(g_keys, g_counts, n_shards, n_q, k, out_keys, out_counts,
                                     torch.cuda.current_stream().cuda_stream, k_in=k_in, d_incomplete=incomplete)
        return cls(local_search, merge, device, group, set_row_order=index.set_sorted_prefix)

    def share_seeds(self, index: "engine.DeviceIndex"):
        """Collective, once after the shards are built: adds up the per-term impact histograms of all shards (one
        all-reduce of n_terms x 256 counters) and gives every shard seed tables of the WHOLE collection, so that its
        searches start from a bound of the global k-th score instead of its own."""
        if self.world == 1:
            return
        torch = self.torch
        n_terms = index.info()["n_terms"]
        hist = torch.zeros((n_terms, 256), dtype=torch.int32, device=self.device)
        stream = torch.cuda.current_stream().cuda_stream
        index.export_seed_hist(hist, stream)
        self.dist.all_reduce(hist, op=self.dist.ReduceOp.SUM, group=self.group)
        index.import_seed_hist(hist, stream)

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape))
        if key not in self._buffers:
            self._buffers[key] = self.torch.zeros(shape, dtype=dtype, device=self.device)
        return self._buffers[key]

    def _flat(self, name, numel, dtype):
        """Flat scratch buffer of at least `numel` elements (grown geometrically, so batches of varying size
        do not accumulate one buffer per size)."""
        buf = self._buffers.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            buf = self.torch.zeros(max(int(numel * 1.5), 1), dtype=dtype, device=self.device)
            self._buffers[name] = buf
        return buf[:numel]

    def _gather_merge(self, keys, counts, n_queries, k_in, k, tag, prove):
        """All-gather rows of k_in keys (keys [Q, k_in] contiguous, counts [Q]) and merge them into the global
        top-k. Returns (keys [Q, k], counts [Q], incomplete [Q] or None)."""
        torch, dist = self.torch, self.dist
        g_keys = self._flat(tag + "g_keys", self.world * n_queries * k_in, torch.int64)
        g_counts = self._flat(tag + "g_counts", self.world * n_queries, torch.int32)
        # concatenated-along-dim-0 form: accepted by both the NCCL and the gloo backend
        dist.all_gather_into_tensor(g_keys.view(self.world * n_queries, k_in), keys, group=self.group)
        dist.all_gather_into_tensor(g_counts, counts, group=self.group)
        out_keys = self._flat(tag + "out_keys", n_queries * k, torch.int64).view(n_queries, k)
        out_counts = self._flat(tag + "out_counts", n_queries, torch.int32)
        incomplete = self._flat(tag + "incomplete", n_queries, torch.int32)
        self.merge(g_keys.view(self.world, n_queries, k_in), g_counts.view(self.world, n_queries), self.world, n_queries,
                   k_in, k, out_keys, out_counts, incomplete)
        return out_keys, out_counts, incomplete if prove else None

    # ---- fused exchange over peer memory (CUDA only) --------------------------------------------------
    def peer_exchange_available(self) -> bool:
        """Collective on first use (every rank must call it at the same point): tries to set the peer buffers up; when
        that fails anywhere (no CUDA IPC / peer access), every rank falls back to the all-gather form for good."""
        if self.world <= 1 or getattr(self.device, "type", "cpu") != "cuda" or os.environ.get("DI_B200_NO_PEER") == "1":
            return False
        if self._peer_ok is None:
            try:
                self._peer = PeerExchange(self.device, self.rank, self.world, 1024, 64, self.group)
                self._second = self.torch.zeros(1, dtype=self.torch.int32, device=self.device)
                self._peer_ok = True
            except RuntimeError as e:
                import warnings
                warnings.warn(f"{e}; using the all-gather exchange")
                self._peer, self._peer_ok = None, False
        return self._peer_ok

    def search_partitioned(self, d_q_terms, d_q_offsets, n_queries: int, max_len: int, k: int):
        """Fused form: returns ((q_lo, q_hi), keys [q_hi - q_lo, k] int64, counts [q_hi - q_lo] int32) — the GLOBAL
        top-k of THIS rank's slice of the queries (`shard_range(n_queries, world, rank)`); the slices of all ranks
        tile the batch. No collective and no host synchronisation in the call. The returned tensors are owned by
        the searcher and overwritten by the next call."""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream
        if self._peer is None or self._peer.q_cap < n_queries or self._peer.k != k:    # collective: same sizes on every rank
            if self._peer is not None:
                self._peer.close()
            self._peer = PeerExchange(self.device, self.rank, self.world, n_queries + n_queries // 4 + 1, k, self.group)
            self._second = torch.zeros(1, dtype=torch.int32, device=self.device)
        ex = self._peer
        rows, counts, row_table, cnt_table = ex.next_set()
        k_in = min(k, self.rows_per_shard(k)) if self.rows_per_shard else shard_k(k, self.world)
        if self.set_row_order:      # the merge reads k_in sorted columns, or (k_in == k) re-selects from whole rows anyway
            self.set_row_order(k_in if k_in < k else 1)
        self.local_search(d_q_terms, d_q_offsets, n_queries, max_len, k, rows, counts)    # this shard's rows
        ex.barrier(stream)                                   # every shard's rows are written and visible
        q_lo, q_hi = shard_range(n_queries, self.world, self.rank)
        n_own = q_hi - q_lo
        out_keys = self._flat("own_keys", max(n_own, 1) * k, torch.int64)[:n_own * k].view(n_own, k)
        out_counts = self._flat("own_counts", max(n_own, 1), torch.int32)[:n_own]
        self._second.zero_()
        engine.merge_pull_device(row_table, cnt_table, self.world, q_lo, n_own, k, k_in, k, out_keys, out_counts,
                                 stream, d_n_second_pass=self._second)
        self.round2_queries = self._second       # device counter of this rank's slice (read it after a synchronize)
        return (q_lo, q_hi), out_keys, out_counts

    def search_tensors(self, d_q_terms, d_q_offsets, n_queries: int, max_len: int, k: int):
        """Device-level entry: returns (keys [Q,k] int64, counts [Q] int32) tensors holding the GLOBAL top-k.
        The returned tensors are owned by the searcher and overwritten by the next call."""
        torch = self.torch
        self.round2_queries = 0
        if self.peer_exchange_available():
            # fused path + one all-gather of the finished slices (callers that can work on a slice use search_partitioned)
            (q_lo, q_hi), own_keys, own_counts = self.search_partitioned(d_q_terms, d_q_offsets, n_queries, max_len, k)
            per = -(-n_queries // self.world)
            pad_keys = self._flat("pad_keys", per * k, torch.int64).view(per, k)
            pad_counts = self._flat("pad_counts", per, torch.int32)
            pad_counts.zero_()
            pad_keys[:q_hi - q_lo].copy_(own_keys)
            pad_counts[:q_hi - q_lo].copy_(own_counts)
            all_keys = self._flat("all_keys", self.world * per * k, torch.int64)
            all_counts = self._flat("all_counts", self.world * per, torch.int32)
            self.dist.all_gather_into_tensor(all_keys.view(self.world * per, k), pad_keys, group=self.group)
            self.dist.all_gather_into_tensor(all_counts, pad_counts, group=self.group)
            return all_keys.view(self.world * per, k)[:n_queries], all_counts[:n_queries]
        keys = self._flat("keys", n_queries * k, torch.int64).view(n_queries, k)
        counts = self._flat("counts", n_queries, torch.int32)
        k_in = min(k, self.rows_per_shard(k)) if self.rows_per_shard else shard_k(k, self.world)
        if self.set_row_order:
            self.set_row_order(0 if self.world == 1 else (k_in if k_in < k else 1))
        self.local_search(d_q_terms, d_q_offsets, n_queries, max_len, k, keys, counts)   # this shard's top-k rows
        if self.world == 1:
            return keys, counts
        if k_in == k:
            out_keys, out_counts, _ = self._gather_merge(keys, counts, n_queries, k, k, "r1_", prove=False)
            return out_keys, out_counts
        head = self._flat("head", n_queries * k_in, torch.int64).view(n_queries, k_in)
        head.copy_(keys[:, :k_in])
        head_counts = self._flat("head_counts", n_queries, torch.int32)
        torch.clamp(counts, max=k_in, out=head_counts)
        out_keys, out_counts, incomplete = self._gather_merge(head, head_counts, n_queries, k_in, k, "r1_", prove=True)
        # the flags derive from gathered data, identical on every rank: all ranks take the same branch
        redo = torch.nonzero(incomplete).flatten()
        if redo.numel():
            n_redo = int(redo.numel())
            self.round2_queries = n_redo
            rows = self._flat("rows", n_redo * k, torch.int64).view(n_redo, k)
            torch.index_select(keys, 0, redo, out=rows)
            row_counts = self._flat("row_counts", n_redo, torch.int32)
            torch.index_select(counts, 0, redo, out=row_counts)
            k2, c2, _ = self._gather_merge(rows, row_counts, n_redo, k, k, "r2_", prove=False)
            out_keys[redo] = k2
            out_counts[redo] = c2
        return out_keys, out_counts

    def search(self, queries: Sequence[Sequence[int]], k: int):
        """queries: term-id lists (identical on every rank). Returns numpy (docids [Q,k], scores [Q,k], counts [Q])."""
        torch = self.torch
        flat, offs = engine.flatten_queries(queries)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint32)
        d_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).to(self.device)
        d_offs = torch.from_numpy(offs.astype(np.int64)).to(self.device)
        max_len = max((len(q) for q in queries), default=0)
        keys, counts = self.search_tensors(d_flat, d_offs, len(queries), max_len, k)
        if self.device is not None and getattr(self.device, "type", "cpu") == "cuda":
            torch.cuda.synchronize()
        scores, docids = unpack_keys(keys.cpu().numpy().view(np.uint64))
        return docids, scores, counts.cpu().numpy().view(np.uint32)
