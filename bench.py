#!/usr/bin/env python
"""bench.py — QPS of top-1000 search on the MS MARCO-shaped synthetic index (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # the CPU restatement (oracle/) on host cores

One "step" = scoring all 6,980 queries against the 8.8 M-document index and selecting each
query's top-1000. With N > 1 (torchrun, one rank per GPU) documents are split into N contiguous
docid ranges; every rank scores all queries on its shard, per-shard top-k keys are all-gathered
over NCCL and merged (strong scaling: the index is fixed, per-GPU work shrinks with N).

`value`  : queries / s with queries and results resident in HBM (device-timed, max over ranks).
`e2e`    : the same through the host-buffer C-ABI call (di_search): pinned host query buffers
           H2D, kernels, result D2H, all inside the timed region.
`roofline`: algorithmic postings bytes (5 B per posting traversed) / time of the
           score_tile launches (CUDA events on the launching stream, inside the library).
`cpu_baseline`: the oracle's C scorer on a bounded sample of the same queries, all host threads;
           the sample doubles as a full-size bit-exact parity check of the GPU results.

Synthetic data (no network): Zipf(1) term popularity over a 30,522-term vocabulary, 120 draws
per document (unique terms kept), 3-decimal log-normal impacts quantized to 8 bits by K1, queries
of 1+Poisson(5) distinct Zipf terms — generated on the GPU with torch (plumbing, not the path).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

ALGO_BYTES_PER_POSTING = 5          # u32 docid + u8 impact: the reference record (defaults.py:28-35)
CHUNK_DOCS = 1 << 20
DATA_SEED = 20240517
IMPACT_CLIP = 12.0


# ----------------------------------------------------------------------------- synthetic data (torch, on device)
def zipf_tables(vocab_size, torch, dev):
    from improving_learned_index_b200 import synthetic
    cdf, perm = synthetic.zipf_cdf(vocab_size)
    return (torch.from_numpy(cdf).to(dev), torch.from_numpy(perm.astype(np.int64)).to(dev))


IMPACT_SKEW = 0.0   # --impact-skew S: impacts of document d are scaled by 1 / (1 + S * d / N) (a quality-ordered collection)


def gen_chunk(chunk_id, lo, hi, n_docs_total, vocab_size, draws, tables, torch, dev, unique=0):
    """Documents [lo, hi) of chunk `chunk_id` (rows are generated for the whole chunk so that any
    shard split sees identical documents). Returns term ids, float64 impacts and the keep mask, every row
    sorted by term id. unique > 0: a document keeps the first `unique` DISTINCT terms of its `draws` Zipf draws
    (configs[1]: "~120 unique expanded terms/doc"); unique == 0: all distinct terms of the draws (round 1)."""
    cdf, perm = tables
    c0 = chunk_id * CHUNK_DOCS
    n = min(CHUNK_DOCS, n_docs_total - c0)
    g = torch.Generator(device=dev)
    g.manual_seed(DATA_SEED * 1_000_003 + chunk_id)
    u = torch.rand((n, draws), generator=g, device=dev, dtype=torch.float64)
    terms = perm[torch.searchsorted(cdf, u).clamp_(max=vocab_size - 1)]
    del u
    z = torch.randn((n, draws), generator=g, device=dev, dtype=torch.float32)
    m = torch.round(1000.0 * torch.exp(0.75 * z.double())).clamp_(0, IMPACT_CLIP * 1000)
    zero = torch.rand((n, draws), generator=g, device=dev, dtype=torch.float32) < 0.01
    m[zero] = 0
    del z, zero
    if IMPACT_SKEW:
        doc = torch.arange(c0, c0 + n, device=dev, dtype=torch.float64)[:, None]
        m = torch.round(m / (1.0 + IMPACT_SKEW * doc / n_docs_total))
    if unique:
        # first occurrence of every term in DRAW order, then the first `unique` of those
        st, order = torch.sort(terms, dim=1, stable=True)
        first_sorted = torch.ones_like(st, dtype=torch.bool)
        first_sorted[:, 1:] = st[:, 1:] != st[:, :-1]
        first = torch.zeros_like(first_sorted).scatter_(1, order, first_sorted)
        del st, order, first_sorted
        chosen = first & (torch.cumsum(first, dim=1) <= unique)
        del first
        terms = torch.where(chosen, terms, torch.full_like(terms, vocab_size))   # dropped draws sort to the end
        terms, order = torch.sort(terms, dim=1)
        m = torch.gather(m, 1, order)
        del order, chosen
        keep = terms < vocab_size
        width = min(draws, unique)                       # at most `unique` kept columns, all at the front
        terms, m, keep = terms[:, :width].contiguous(), m[:, :width].contiguous(), keep[:, :width].contiguous()
    else:
        terms, order = torch.sort(terms, dim=1)
        m = torch.gather(m, 1, order)
        del order
        keep = torch.ones_like(terms, dtype=torch.bool)
        keep[:, 1:] = terms[:, 1:] != terms[:, :-1]
    sl = slice(lo - c0, hi - c0)
    return terms[sl], (m[sl] / 1000.0), keep[sl]



def build_shard_arrays(doc_lo, doc_hi, n_docs_total, vocab_size, draws, torch, dev, quantize_fn, unique=0):
    """Doc-major arrays of the shard [doc_lo, doc_hi): term ids (u32 as int32 bits), u8 impacts (quantized by
    `quantize_fn`, postings with value 0 dropped as quantize.py:45 does), u64 doc offsets (local docs)."""
    tables = zipf_tables(vocab_size, torch, dev)
    t_parts, v_parts, c_parts = [], [], []
    for chunk_id in range(doc_lo // CHUNK_DOCS, (doc_hi - 1) // CHUNK_DOCS + 1):
        lo = max(doc_lo, chunk_id * CHUNK_DOCS)
        hi = min(doc_hi, (chunk_id + 1) * CHUNK_DOCS)
        terms, impacts, keep = gen_chunk(chunk_id, lo, hi, n_docs_total, vocab_size, draws, tables, torch, dev, unique)
        q = quantize_fn(impacts.reshape(-1).contiguous()).reshape(impacts.shape)
        keep &= q > 0
        t_parts.append(terms[keep].to(torch.int32))
        v_parts.append(q[keep].to(torch.uint8))
        c_parts.append(keep.sum(dim=1))
        del terms, impacts, keep, q
    counts = torch.cat(c_parts)
    offs = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=offs[1:])
    return torch.cat(t_parts), torch.cat(v_parts), offs


def workload_config(args, k):
    name = ("configs[3]: 100K queries of ~6 terms, batches of %d" % args.batch if args.workload == "c4"
            else "configs[1]: MS MARCO passage-shaped synthetic index")
    return {"workload": "%s, Zipf(1) terms, top-%d" % (name, k), "docs": args.docs, "vocab": args.vocab,
            "draws_per_doc": args.draws, "impact_skew": args.impact_skew,
            "unique_terms_per_doc": args.unique_terms or "all distinct terms of the draws (~91 of 120; round-1 workload)",
            "queries": args.queries, "top_k": k}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.stop, self.gpu = [], threading.Event(), gpu_index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[2]) for r in self.rows)}


def committed_profile(n_docs, n_queries, top_k, world, unique_terms):
    """Numbers that only a profiler can give, read from the committed `ncu --set full` summary of the dominant
    kernel (profiles/r2_k3_limiter.json, written by tools/ncu_summary.py) — used only when that capture was taken
    on this very configuration: DRAM bytes per launch and the SM-side utilisation figures that name the limiter."""
    try:
        t = json.load(open(REPO / "profiles" / "r2_k3_limiter.json"))
        c = t["config"]
        if (c["docs"], c["queries"], c["top_k"], c["n_gpus"], c.get("unique_terms", 0)) == (n_docs, n_queries, top_k, world, unique_terms):
            return t
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------- main arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from improving_learned_index_b200 import _native, engine, synthetic
    from improving_learned_index_b200.sharded import ShardedSearcher, shard_range

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    _native.set_device(local)
    dev = torch.device(f"cuda:{local}")
    L = _native.lib()
    stream = torch.cuda.current_stream().cuda_stream
    N, V, k = args.docs, args.vocab, args.top_k
    doc_lo, doc_hi = shard_range(N, world, rank)

    # ---- build: synthetic doc-major lists -> K1 quantize -> K2 invert -> tiled shard
    t0 = time.time()

    def quantize_fn(x):   # K1 on device buffers; max over the collection is the clip value by construction
        out = torch.empty(x.numel(), dtype=torch.int32, device=dev)
        _native.check(L.di_quantize_f64_dev(x.data_ptr(), x.numel(), IMPACT_CLIP, out.data_ptr(), stream))
        return out
    terms, imps, offs = build_shard_arrays(doc_lo, doc_hi, N, V, args.draws, torch, dev, quantize_fn, args.unique_terms)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    P = terms.numel()
    toff = torch.empty(V + 1, dtype=torch.int64, device=dev)
    docids = torch.empty(P, dtype=torch.int32, device=dev)
    vals = torch.empty(P, dtype=torch.uint8, device=dev)
    invert_runs = []
    for _ in range(2):      # the first call also pays for mapping ~13 GB of fresh sort scratch (cudaMalloc)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _native.check(L.di_invert_dev(terms.data_ptr(), imps.data_ptr(), offs.data_ptr(), doc_hi - doc_lo, V, P,
                                      toff.data_ptr(), docids.data_ptr(), vals.data_ptr(), None, stream))
        ev1.record()
        torch.cuda.synchronize()
        invert_runs.append(ev0.elapsed_time(ev1))
    invert_ms = invert_runs[-1]
    build_parity = None
    if args.verify_build and world == 1:              # configs[4]: the whole K2 output vs the CPU oracle, bit for bit
        from oracle import oracle
        t_cpu = time.time()
        o_toff, o_docs, o_vals = oracle.invert(terms.cpu().numpy().view(np.uint32), imps.cpu().numpy(),
                                               offs.cpu().numpy().astype(np.uint64), V)
        t_cpu = time.time() - t_cpu
        same = (np.array_equal(o_toff, toff.cpu().numpy().astype(np.uint64))
                and np.array_equal(o_docs, docids.cpu().numpy().view(np.uint32))
                and np.array_equal(o_vals, vals.cpu().numpy()))
        build_parity = {"postings": int(P), "bit_exact": bool(same), "oracle_invert_s": round(t_cpu, 2),
                        "against": "oracle/di_oracle.c dio_invert (create.py:31-46)"}
        del o_toff, o_docs, o_vals
        if not same:
            raise SystemExit("PARITY FAILURE: GPU inversion differs from the oracle at full size")
    docids += doc_lo                                  # docids stay global across shards
    torch.cuda.synchronize()
    t1 = time.time()
    index_flags = _native.INDEX_TILE_BOUNDS if args.tile_bounds else 0
    if args.index_from == "docmajor":                 # one segmented two-pass sort, no term-major detour
        index = engine.DeviceIndex.from_docmajor_device(terms, imps, offs, doc_hi - doc_lo, V, P, doc_lo=doc_lo,
                                                        tile_docs=args.tile_docs, dense_ratio=args.dense_ratio,
                                                        cand_slack=args.cand_slack, flags=index_flags)
    else:                                             # from the inverted CSR (the reference's index format)
        index = engine.DeviceIndex.from_csr_device(toff, docids, vals, V, P, doc_lo=doc_lo, doc_hi=max(doc_hi, doc_lo + 1),
                                                   tile_docs=args.tile_docs, dense_ratio=args.dense_ratio,
                                                   cand_slack=args.cand_slack, flags=index_flags)
    t_tile = time.time() - t1
    del terms, imps, offs
    info = index.info()
    build_info = {"generate_s": round(t_gen, 2), "invert_ms": round(invert_ms, 1), "invert_ms_first_call": round(invert_runs[0], 1),
                  "tile_layout_s": round(t_tile, 3), "index_from": args.index_from, "invert_postings_per_s": round(P / (invert_ms * 1e-3)),
                  "invert_gbs_at_17B": round(17 * P / (invert_ms * 1e-3) / 1e9, 1)}
    n_idx = torch.tensor([info["n_postings"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(n_idx, op=dist.ReduceOp.SUM)
    total_postings_index = int(n_idx.item())

    # ---- queries
    queries = synthetic.make_queries(args.queries, vocab_size=V, seed=7)
    flat, qoffs = engine.flatten_queries(queries)
    max_len = max(len(q) for q in queries)
    Q = len(queries)
    h_flat = torch.from_numpy(flat.astype(np.int64)).to(torch.int32).pin_memory()
    h_offs = torch.from_numpy(qoffs.astype(np.int64)).pin_memory()
    d_flat, d_offs = h_flat.to(dev), h_offs.to(dev)
    df = index.term_df(flat)
    local_postings = int(df.sum())
    searcher = ShardedSearcher.for_device_index(index, dev)      # local rows -> stream barrier -> fused pull-merge
    if world > 1 and not args.local_seeds:
        searcher.share_seeds(index)           # one all-reduce at build time: every shard seeds from the whole collection
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    B = args.batch or Q                       # queries per search call (configs[3] uses batches of 4096)
    batches = [(q0, min(q0 + B, Q)) for q0 in range(0, Q, B)]

    fused = searcher.peer_exchange_available()    # N > 1: exchange + merge as one kernel over peer memory, no collective

    def search_batch(q0, q1):
        """-> ((first, last) query of this rank's result rows, keys, counts)"""
        if fused:
            (lo, hi), keys, counts = searcher.search_partitioned(d_flat, d_offs[q0:q1 + 1], q1 - q0, max_len, k)
            return (q0 + lo, q0 + hi), keys, counts
        keys, counts = searcher.search_tensors(d_flat, d_offs[q0:q1 + 1], q1 - q0, max_len, k)
        return ((q0, q1) if rank == 0 else (q0, q0)), keys, counts   # all-gather form: rank 0 reports the rows

    def step_device():
        out = None
        for q0, q1 in batches:
            out = search_batch(q0, q1)
        return out

    h_docs = torch.empty((Q, k), dtype=torch.int32).pin_memory()
    h_scores = torch.empty((Q, k), dtype=torch.int32).pin_memory()
    h_counts = torch.empty(Q, dtype=torch.int32).pin_memory()
    if world > 1:
        dd = torch.empty((B, k), dtype=torch.int32, device=dev)
        ds = torch.empty((B, k), dtype=torch.int32, device=dev)

    def step_e2e():
        for q0, q1 in batches:
            if world == 1:      # H2D + kernels + D2H inside the C-ABI call
                index.search_flat(h_flat, h_offs[q0:q1 + 1], k, h_docs[q0:q1], h_scores[q0:q1], h_counts[q0:q1])
            else:               # every rank: this batch's queries H2D, search + exchange, its own slice of the results D2H
                t0, t1 = int(h_offs[q0]), int(h_offs[q1])
                d_flat[t0:t1].copy_(h_flat[t0:t1], non_blocking=True)
                d_offs[q0:q1 + 1].copy_(h_offs[q0:q1 + 1], non_blocking=True)
                (r0, r1), m_keys, m_counts = search_batch(q0, q1)
                n = r1 - r0
                if n:
                    engine.unpack_keys_device(m_keys, n * k, dd, ds, stream)
                    h_docs[r0:r1].copy_(dd[:n], non_blocking=True)
                    h_scores[r0:r1].copy_(ds[:n], non_blocking=True)
                    h_counts[r0:r1].copy_(m_counts[:n], non_blocking=True)
                torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    index.timings()                          # the library accumulates until read: drop the warm-up records
    # ---- timed: device-resident
    score_ms, final_ms, step_ms = [], [], []
    # (rank 0 samples the clocks: eight ranks forking nvidia-smi ten times a second starve the host threads that drive the GPUs)
    with (ClockSampler(local) if rank == 0 else contextlib.nullcontext()) as clocks:
        barrier()
        for _ in range(args.steps):
            flush.fill_(1)                      # L2 flush between iterations (outside the event pair)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_device()
            b.record()
            b.synchronize()
            step_ms.append(a.elapsed_time(b))
            t = index.timings()
            score_ms.append(t["score_ms"])
            final_ms.append(t["finalize_ms"])
            launches = t                     # kernels this rank's library launched in this step
            # + the exchange per batch: fused = stream barrier + pull-merge pass 1 (+ pass 2 when rows are cut, world > 2);
            #   fallback = merge_gather + finalize + merge_check
            n_launches = t["score_launches"] + t["other_launches"] + (((3 if world > 2 else 2) if fused else 3) * len(batches) if world > 1 else 0)
        barrier()
        # ---- timed: end to end through the host-buffer call
        step_e2e()
        e2e_ms = []
        for _ in range(args.steps):
            flush.fill_(1)
            barrier()
            t_a = time.perf_counter()
            step_e2e()
            e2e_ms.append((time.perf_counter() - t_a) * 1e3)
        barrier()
    index.timings()                          # drop the e2e calls' records
    total_ms = torch.tensor([sum(step_ms), sum(e2e_ms), sum(score_ms)], dtype=torch.float64, device=dev)
    post = torch.tensor([local_postings], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(post, op=dist.ReduceOp.SUM)
    total_dev_ms, total_e2e_ms, total_score_ms = total_ms.tolist()
    total_postings = int(post.item())

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(REPO / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # per-GPU roofline of the dominant kernel (score_persistent_kernel): this rank's ALGORITHMIC postings bytes
        # (5 B per posting traversed) / its launch time. The kernel serves every posting of a tile to the
        # whole batch from L2 and stores hot lists as one byte per document, so `frac` can exceed 1: it says how much
        # HBM traffic a one-query-at-a-time streaming scorer would need, not how busy the DRAM is. `dram_frac` is the
        # measured DRAM traffic / time / peak and `limiter` names what actually bounds the kernel (ncu, profiles/).
        score_s = float(np.mean(score_ms)) * 1e-3
        achieved = ALGO_BYTES_PER_POSTING * local_postings / score_s / 1e9
        prof = committed_profile(N, Q, k, world, args.unique_terms)
        traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"] if prof else None
        cfg = workload_config(args, k)
        cfg.update({"postings": total_postings_index, "batch": B, "sharding": f"docid-range x{world}",
                    "tile_docs": info["tile_docs"],
                    "l2": "index payload (%.1f GB/GPU) exceeds L2 and a 256 MB buffer is written between timed steps"
                          % (info["payload_bytes"] / 1e9)})
        out = {
            "metric": "QPS top-%d on 8.8M-doc MS MARCO-shaped index" % k,
            "value": round(Q * args.steps / (total_dev_ms * 1e-3), 2), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(total_dev_ms / args.steps, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 impacts, u16/int32 accumulators", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": round(Q * args.steps / (total_e2e_ms * 1e-3), 2), "unit": "queries/s",
                    "h2d_bytes_per_step": int(h_flat.numel() * 4 + h_offs.numel() * 8),
                    "d2h_bytes_per_step": int(Q * k * 8 + Q * 4)},
            # kernels of this repo launched inside the `value` timed region, all ranks
            "gpu_launches": int(n_launches * args.steps * world),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic,
                         "frac_is": "algorithmic postings bytes / time / peak (SURVEY 8d); > 1 = traffic the L2-resident "
                                    "batch and the byte-per-document lists do not send to HBM",
                         "dram_frac": round(traffic / score_s / 1e9 / peak, 4) if traffic else None,
                         "limiter": prof.get("limiter") if prof else None,
                         "limiter_frac": prof.get("limiter_frac") if prof else None,
                         "sm": prof.get("sm") if prof else None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                         "kernel": "score_persistent_kernel", "launches_per_step": launches["score_launches"],
                         "algorithmic_bytes_per_step_this_gpu": ALGO_BYTES_PER_POSTING * local_postings,
                         "score_ms_per_step": round(float(np.mean(score_ms)), 3),
                         "finalize_ms_per_step": round(float(np.mean(final_ms)), 3),
                         "tile_lanes": launches["lanes"], "tile_bounds": bool(args.tile_bounds),
                         "tiles_skipped_per_step": int(launches.get("tiles_skipped", 0))},
            "clocks": clocks.summary(),
            "index": {"postings_this_gpu": info["n_postings"], "payload_gb": round(info["payload_bytes"] / 1e9, 3),
                      "dense_segments": info["n_dense_segments"], "sparse_segments": info["n_sparse_segments"],
                      "dense_posting_frac": round(info["n_dense_postings"] / max(info["n_postings"], 1), 3),
                      "tiles": info["n_tiles"]},
            "build": build_info,
            "postings_per_query": round(total_postings / Q),
            "build_parity": build_parity,
            "shared_seeds": bool(world > 1 and not args.local_seeds),
            "exchange": ("fused peer-memory pull (CUDA IPC + stream barrier + di_merge_pull_dev), queries partitioned over ranks"
                         if fused else ("NCCL all-gather + K5" if world > 1 else None)),
            "second_pass_queries_rank0_last_step": int(searcher.round2_queries),
        }
        if world == 1 and args.cpu_sample > 0:
            out["cpu_baseline"], out["parity"] = cpu_baseline_and_parity(
                args, toff, docids, vals, queries, (h_docs, h_scores, h_counts), torch)
        if world == 1 and args.file_legs and args.workload == "c2" and args.cpu_sample > 0:
            legs = file_based_legs(args, toff, docids, vals, queries, (h_docs, h_scores, h_counts), torch, dev, L, stream)
            for key in ("cpu_baseline_python", "api_e2e", "load"):
                if key in legs:
                    out[key] = legs[key]
            if "load" in legs:
                out["build"]["load"] = legs["load"]
            if "error" in legs or "unavailable" in legs:
                out["file_legs_note"] = legs.get("error") or legs.get("unavailable")
            if "api_e2e" in legs:
                out["api_e2e"]["vs_e2e"] = {key: round(v["value"] / out["e2e"]["value"], 3) for key, v in legs["api_e2e"].items()}
    if world > 1 and args.verify_sharded:
        # every rank's merged result vs ONE index over all documents built on rank 0's GPU (itself checked against
        # the oracle in the 1-GPU run): the sharded path must be bit-identical
        m_keys, m_counts = searcher.search_tensors(d_flat, d_offs, Q, max_len, k)
        m_keys, m_counts = m_keys.clone(), m_counts.clone()
        same = None
        if rank == 0:
            t_all, v_all, o_all = build_shard_arrays(0, N, N, V, args.draws, torch, dev, quantize_fn, args.unique_terms)
            P_all = t_all.numel()
            toff_a = torch.empty(V + 1, dtype=torch.int64, device=dev)
            docs_a = torch.empty(P_all, dtype=torch.int32, device=dev)
            vals_a = torch.empty(P_all, dtype=torch.uint8, device=dev)
            _native.check(L.di_invert_dev(t_all.data_ptr(), v_all.data_ptr(), o_all.data_ptr(), N, V, P_all,
                                          toff_a.data_ptr(), docs_a.data_ptr(), vals_a.data_ptr(), None, stream))
            del t_all, v_all, o_all
            whole = engine.DeviceIndex.from_csr_device(toff_a, docs_a, vals_a, V, P_all, doc_lo=0, doc_hi=N)
            w_keys = torch.zeros((Q, k), dtype=torch.int64, device=dev)
            w_counts = torch.zeros(Q, dtype=torch.int32, device=dev)
            whole.search_device(d_flat, d_offs, Q, max_len, k, w_keys, w_counts, stream)
            torch.cuda.synchronize()
            valid = torch.arange(k, device=dev)[None, :] < w_counts[:, None]
            same = bool(torch.equal(w_counts, m_counts) and torch.equal(w_keys[valid], m_keys[valid]))
            whole.close()
            out["sharded_parity"] = {"queries_checked": Q, "bit_exact": same,
                                     "against": "one index over all documents on rank 0 (same kernels, 1 shard)"}
        dist.barrier()
        if same is False:
            raise SystemExit("PARITY FAILURE: the sharded result differs from the single-index result")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def cpu_baseline_and_parity(args, toff, docids, vals, queries, timed_result, torch):
    """Oracle (C port, all host threads) on a bounded query sample of the same index. The GPU rows compared with
    it are taken FROM THE OUTPUT OF THE LAST TIMED END-TO-END STEP (the full batch through the host-buffer
    C-ABI call), so the parity claim covers exactly the code path that produced the number."""
    from oracle import oracle
    h_toff = toff.cpu().numpy().astype(np.uint64)
    h_docs = docids.cpu().numpy().view(np.uint32)
    h_vals = vals.cpu().numpy()
    rng = np.random.default_rng(123)
    pick = sorted(rng.choice(len(queries), size=min(args.cpu_sample, len(queries)), replace=False).tolist())
    sample = [queries[i] for i in pick]
    threads = host_threads()
    t0 = time.perf_counter()
    o_docs, o_scores, o_counts, o_post = oracle.score_topk_csr(h_toff, h_docs, h_vals, args.docs, sample, args.top_k,
                                                              n_threads=threads)
    dt = time.perf_counter() - t0
    g_docs, g_scores, g_counts = (t.numpy() for t in timed_result)
    ok = True
    for i, qi in enumerate(pick):
        n = int(o_counts[i])
        ok = (ok and int(g_counts[qi]) == n and np.array_equal(g_docs[qi, :n].view(np.uint32), o_docs[i, :n])
              and np.array_equal(g_scores[qi, :n], o_scores[i, :n]))
    base = {"value": round(len(sample) / dt, 3), "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{len(sample)} of the {len(queries)} queries (seed 123), full index, oracle/di_oracle.c with OpenMP",
            "postings_per_s": round(float(o_post.sum()) / dt), "seconds": round(dt, 2)}
    parity = {"queries_checked": len(sample), "bit_exact": bool(ok), "path": "timed batch",
              "rows_from": "output buffers of the last timed end-to-end step (%d queries per call)" % (args.batch or len(queries)),
              "against": "oracle/di_oracle.c (pinned to reference golden vectors)"}
    if not ok:
        raise SystemExit("PARITY FAILURE at full size: GPU results of the timed batch differ from the oracle")
    return base, parity


def file_based_legs(args, toff, docids, vals, queries, timed_result, torch, dev, L, stream):
    """Everything that needs the index as the reference's three FILES (rank 0, N = 1, configs[1] only): the files are
    written once from this run's CSR (di_serialize: byte format of create.py:27-51) into a RAM-backed scratch directory,
    then   (1) `cpu_baseline_python`: the reference's own unmodified Python reader, timed on them
               (tools/py_reference_timing.py);
           (2) `load`: InvertedIndex(dir) — the reference-shaped loader: file images -> pinned chunks -> HBM -> tiled index;
           (3) `api_e2e`: the drop-in classes a user of the reference calls — Ranker.run (query TSV -> run file) and
               InvertedIndex.score_batch — on all queries, wall clock, against `e2e` (the array-level C-ABI call)."""
    import shutil
    import tempfile
    out = {}
    P, V, Q, k = docids.numel(), args.vocab, len(queries), args.top_k
    need = 5 * P + 16 * V + 40 * Q * k + (1 << 20)
    root = next((d for d in ("/dev/shm", tempfile.gettempdir()) if os.path.isdir(d) and shutil.disk_usage(d).free > need + (2 << 30)), None)
    if root is None:
        return {"unavailable": "no scratch space for the %.1f GB .dat file" % (need / 1e9)}
    tmp = Path(tempfile.mkdtemp(prefix="di_ref_index_", dir=root))
    try:
        from improving_learned_index_b200 import InvertedIndex, _native, synthetic
        from improving_learned_index_b200.evaluation import Ranker
        t0 = time.perf_counter()
        d_dat = torch.empty(5 * P, dtype=torch.uint8, device=dev)
        d_idx = torch.empty(2 * V, dtype=torch.int64, device=dev)
        _native.check(L.di_serialize_dev(toff.data_ptr(), docids.data_ptr(), vals.data_ptr(), V, P, d_dat.data_ptr(),
                                         d_idx.data_ptr(), stream))
        torch.cuda.synchronize()
        d_dat.cpu().numpy().tofile(tmp / "inverted_index.dat")
        d_idx.cpu().numpy().tofile(tmp / "inverted_index.idx")
        del d_dat, d_idx
        (tmp / "vocab.txt").write_text(''.join(synthetic.term_name(t) + '\n' for t in range(V)))
        out["index_files"] = {"dat_gb": round(5 * P / 1e9, 3), "write_s": round(time.perf_counter() - t0, 2), "dir": root}
        # (1) the reference's own Python
        if args.py_ref_seconds > 0:
            if not (REPO / "baseline" / "_ref" / "src").is_dir():
                out["cpu_baseline_python"] = {"unavailable": "baseline/_ref/src missing: __graft_entry__.build() stages it where /root/reference exists"}
            else:
                env = dict(os.environ)
                env.pop("OMP_NUM_THREADS", None)
                r = subprocess.run([sys.executable, str(REPO / "tools" / "py_reference_timing.py"), "--index-dir", str(tmp),
                                    "--queries", str(args.queries), "--vocab", str(V), "--seconds", str(args.py_ref_seconds)],
                                   capture_output=True, text=True, timeout=60 + 12 * args.py_ref_seconds, env=env)
                out["cpu_baseline_python"] = (json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else
                                              {"unavailable": "py_reference_timing.py failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200]})
        # (2) loader
        t0 = time.perf_counter()
        index = InvertedIndex(tmp)
        torch.cuda.synchronize()
        load_s = time.perf_counter() - t0
        out["load"] = {"seconds": round(load_s, 3), "dat_gb_per_s": round(5 * P / 1e9 / load_s, 2),
                       "what": "InvertedIndex(index_dir): vocab.txt + .idx + mmap'ed .dat -> pinned chunks -> HBM -> decode -> tiled index"}
        # (3) the drop-in API
        qfile, run = tmp / "queries.tsv", tmp / "run.tsv"
        qfile.write_text(''.join("%d\t%s\n" % (i, ' '.join(synthetic.term_name(t) for t in q)) for i, q in enumerate(queries)))
        ranker = Ranker(index, qfile, run, query_processor=lambda text: text.split(), top_k=k, batch_size=args.api_batch)
        times = []
        for _ in range(3):                                  # first pass = warm-up (workspace allocation)
            if run.exists():
                run.unlink()
            t0 = time.perf_counter()
            ranker.run()
            times.append(time.perf_counter() - t0)
        g_docs, g_scores, g_counts = (t.numpy() for t in timed_result)
        same, n_rows = True, 0
        with open(run) as f:                                # the rows of the first queries against the timed C-ABI batch
            for line in f:
                qid, pid, rank, score = line.split('\t')
                qi, r = int(qid), int(rank) - 1
                if qi >= 50:
                    break
                n_rows += 1
                same = same and r < g_counts[qi] and int(pid) == int(g_docs[qi, r].view(np.uint32)) and int(score) == int(g_scores[qi, r])
        term_lists = [[synthetic.term_name(t) for t in q] for q in queries]
        sb = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = index.score_batch(term_lists, top_k=k)
            sb.append(time.perf_counter() - t0)
        same_sb = bool(np.array_equal(res.counts, g_counts.view(np.uint32)) and np.array_equal(res.docids[:64], g_docs[:64].view(np.uint32)))
        out["api_e2e"] = {"ranker_run": {"value": round(Q / min(times[1:]), 1), "unit": "queries/s", "seconds": round(min(times[1:]), 4),
                                         "what": "Ranker.run(): %d queries from a TSV -> term ids -> GPU search in batches of %d -> "
                                                 "run file of %d rows (%.0f MB)" % (Q, args.api_batch, int(g_counts.sum()), run.stat().st_size / 1e6),
                                         "rows_checked_against_timed_batch": n_rows, "rows_match": bool(same)},
                          "score_batch": {"value": round(Q / min(sb[1:]), 1), "unit": "queries/s", "seconds": round(min(sb[1:]), 4),
                                          "what": "InvertedIndex.score_batch(list of term-string lists, top_k): arrays behind a lazy list view",
                                          "rows_match": same_sb}}
        if not (same and same_sb):
            raise SystemExit("PARITY FAILURE: the drop-in API returned other rows than the timed C-ABI batch")
        index.device_index.close()
        return out
    except SystemExit:
        raise
    except Exception as e:          # a reported baseline must never take the GPU measurement down with it
        out["error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def host_threads():
    """Host threads the CPU legs use: every core this process may run on — NOT OMP_NUM_THREADS, which torchrun
    sets to 1 for its workers."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's algorithm on the host cores: oracle/di_oracle.c (a C restatement of
    inverted_index.py:55-62; the reference's own Python cannot travel to the GPU box), all threads.
    Index data comes from the same seeded generator; quantize + inversion also run on the CPU oracle."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    from improving_learned_index_b200 import synthetic
    from oracle import oracle
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    N, V, k = args.docs, args.vocab, args.top_k
    scale_max = IMPACT_CLIP

    def quantize_fn(x):      # CPU oracle arithmetic (quantize.py:13-14), not the CUDA kernel
        q = oracle.quantize(x.cpu().numpy(), scale_max)
        return torch.from_numpy(q.astype(np.int32)).to(x.device)
    terms, imps, offs = build_shard_arrays(0, N, N, V, args.draws, torch, dev, quantize_fn, args.unique_terms)
    toff, docids, vals = oracle.invert(terms.cpu().numpy().view(np.uint32), imps.cpu().numpy(),
                                       offs.cpu().numpy().astype(np.uint64), V)
    del terms, imps, offs
    queries = synthetic.make_queries(args.queries, vocab_size=V, seed=7)
    threads = host_threads()           # explicit: under torchrun OMP_NUM_THREADS is 1
    per_step = max(1, min(args.ref_queries_per_step, len(queries)))
    rng = np.random.default_rng(123)
    times, n_done = [], 0
    for s in range(args.warmup + args.steps):
        pick = rng.choice(len(queries), size=per_step, replace=False)
        sample = [queries[i] for i in pick]
        t0 = time.perf_counter()
        oracle.score_topk_csr(toff, docids, vals, N, sample, k, n_threads=threads)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            n_done += per_step
    qps = n_done / sum(times)
    out = {"impl": "reference", "metric": "QPS top-%d on 8.8M-doc MS MARCO-shaped index" % k, "value": round(qps, 3),
           "unit": "queries/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(1e3 * sum(times) / args.steps, 2), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "u8 impacts, int32 accumulators", "data": "synthetic",
           "config": dict(workload_config(args, k), queries_per_step=per_step, postings=int(docids.size)),
           "cpu_baseline": {"value": round(qps, 3), "unit": "queries/s", "cores": threads, "kind": "port",
                            "sample": f"{per_step} random queries of the {len(queries)} per step, full 8.8M-doc index"},
           "e2e": {"value": round(qps, 3), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=8_841_823)
    ap.add_argument("--vocab", type=int, default=30522)
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2 = BASELINE configs[1]/[2] (6,980 queries, top-1000, one batch); "
                         "c4 = configs[3] (100,000 queries of ~6 terms, batches of 4,096, top-100)")
    ap.add_argument("--unique-terms", type=int, default=120,
                    help="distinct terms kept per document (configs[1]: ~120); 0 = round-1 workload: all distinct "
                         "terms of --draws 120 draws (~91 per document)")
    ap.add_argument("--draws", type=int, default=0, help="Zipf draws per document (0 = 208 with --unique-terms, else 120)")
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--top-k", type=int, default=0)
    ap.add_argument("--index-from", default="docmajor", choices=["docmajor", "csr"],
                    help="build the device index straight from the doc-major collection, or from the inverted CSR")
    ap.add_argument("--impact-skew", type=float, default=0.0,
                    help="S > 0: impacts of document d are divided by 1 + S*d/N — a quality-ordered collection, the case "
                         "exact tile skipping (--tile-bounds) is for")
    ap.add_argument("--local-seeds", action="store_true", help="N > 1: keep per-shard seed tables (A/B of share_seeds)")
    ap.add_argument("--tile-bounds", action="store_true", help="build the index with DI_INDEX_TILE_BOUNDS (exact tile skipping)")
    ap.add_argument("--tile-docs", type=int, default=0)
    ap.add_argument("--dense-ratio", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0, help="queries per search call (0 = all queries at once)")
    ap.add_argument("--cand-slack", type=int, default=0, help="candidate slots kept per query between tiles (0 = default)")
    ap.add_argument("--verify-build", action="store_true", help="compare the full GPU inversion with the CPU oracle (~1 min)")
    ap.add_argument("--verify-sharded", action="store_true",
                    help="N > 1: compare the merged result of every query with a single index built on rank 0")
    ap.add_argument("--cpu-sample", type=int, default=64, help="queries in the timed CPU baseline / parity sample")
    ap.add_argument("--ref-queries-per-step", type=int, default=32)
    ap.add_argument("--no-file-legs", dest="file_legs", action="store_false",
                    help="skip the legs that need the index as files (Python reference, loader, drop-in API throughput)")
    ap.add_argument("--api-batch", type=int, default=1745, help="queries per GPU call of Ranker.run in the api_e2e leg")
    ap.add_argument("--py-ref-seconds", type=float, default=10.0,
                    help="time budget per leg of the unmodified Python reference (single process and Pool); 0 = skip")
    args = ap.parse_args()
    if args.draws == 0:
        args.draws = 208 if args.unique_terms else 120
    c4 = args.workload == "c4"
    args.queries = args.queries or (100_000 if c4 else 6980)
    args.top_k = args.top_k or (100 if c4 else 1000)
    if c4 and not args.batch:
        args.batch = 4096
    global IMPACT_SKEW
    IMPACT_SKEW = args.impact_skew
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        if int(os.environ.get("WORLD_SIZE", 1)) != args.gpus and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
        run_b200(args)


if __name__ == "__main__":
    main()
