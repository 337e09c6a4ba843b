"""Docid-range sharding across the GPUs of one box (one process per GPU, torch.distributed).

New functionality: the reference searches on one host only (its parallelism is a per-query
process pool, ranker.py:44-46). Documents are split into `world_size` contiguous docid ranges;
rank r holds a full-vocabulary shard of range r with GLOBAL docids. Every rank scores every
query on its shard (no data-path collective: a document's score depends only on its own
postings), then the per-shard top-k keys — (score << 32 | ~docid), sorted — are all-gathered
(NCCL over NVLink on GPUs; gloo in the CPU tests) and merged by K5 into the global top-k, which
is identical to the single-GPU result because the key order is total.

The collective and the two compute steps are injected, so the plumbing (ranges, tensor layout of
the gather, count handling) is testable on CPU with world_size 2 while the CUDA path plugs in
DeviceIndex.search_device and engine.merge_topk_device.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from . import engine


def shard_range(n_docs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous docid range [lo, hi) of `rank`: ceil(n_docs / world_size) docs per shard."""
    per = -(-n_docs // world_size)
    return min(rank * per, n_docs), min((rank + 1) * per, n_docs)


def pack_keys(scores: np.ndarray, docids: np.ndarray) -> np.ndarray:
    """(score, docid) -> ranking key; larger key = better, equal scores order by ascending docid."""
    return (scores.astype(np.uint64) << np.uint64(32)) | (~docids.astype(np.uint32)).astype(np.uint64)


def unpack_keys(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    k = keys.astype(np.uint64)
    return (k >> np.uint64(32)).astype(np.int32), ~(k & np.uint64(0xFFFFFFFF)).astype(np.uint32)


class ShardedSearcher:
    """search(): local top-k on this rank's shard -> all_gather -> merge. Call on every rank with the
    same queries; every rank gets the global result.

    local_search(d_q_terms, d_q_offsets, n_queries, max_len, k, out_keys, out_counts) fills this shard's
    sorted keys [Q, k] (int64 bit patterns) and counts [Q] (int32) — torch tensors on `device`.
    merge(gathered_keys [G,Q,k], gathered_counts [G,Q], G, Q, k, out_keys, out_counts) writes the global top-k.
    """

    def __init__(self, local_search: Callable, merge: Callable, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.local_search, self.merge = local_search, merge
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._buffers = None

    @classmethod
    def for_device_index(cls, index: "engine.DeviceIndex", device, group=None) -> "ShardedSearcher":
        import torch

        def local_search(qt, qo, n_q, max_len, k, out_keys, out_counts):
            index.search_device(qt, qo, n_q, max_len, k, out_keys, out_counts, torch.cuda.current_stream().cuda_stream)

        def merge(g_keys, g_counts, n_shards, n_q, k, out_keys, out_counts):
            engine.merge_topk_device(g_keys, g_counts, n_shards, n_q, k, out_keys, out_counts,
                                     torch.cuda.current_stream().cuda_stream)
        return cls(local_search, merge, device, group)

    def search_tensors(self, d_q_terms, d_q_offsets, n_queries: int, max_len: int, k: int):
        """Device-level entry: returns (keys [Q,k] int64, counts [Q] int32) tensors holding the GLOBAL top-k.
        The returned tensors are owned by the searcher and overwritten by the next call."""
        torch, dist = self.torch, self.dist
        if self._buffers is None or self._buffers[0] != (n_queries, k):      # reused across calls of one shape
            def buf(*shape, dtype):
                return torch.zeros(shape, dtype=dtype, device=self.device)
            self._buffers = ((n_queries, k), buf(n_queries, k, dtype=torch.int64), buf(n_queries, dtype=torch.int32),
                             buf(self.world, n_queries, k, dtype=torch.int64), buf(self.world, n_queries, dtype=torch.int32),
                             buf(n_queries, k, dtype=torch.int64), buf(n_queries, dtype=torch.int32))
        _, keys, counts, g_keys, g_counts, out_keys, out_counts = self._buffers
        self.local_search(d_q_terms, d_q_offsets, n_queries, max_len, k, keys, counts)
        if self.world == 1:
            return keys, counts
        # concatenated-along-dim-0 form: accepted by both the NCCL and the gloo backend
        dist.all_gather_into_tensor(g_keys.view(self.world * n_queries, k), keys, group=self.group)
        dist.all_gather_into_tensor(g_counts.view(self.world * n_queries), counts, group=self.group)
        self.merge(g_keys, g_counts, self.world, n_queries, k, out_keys, out_counts)
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
