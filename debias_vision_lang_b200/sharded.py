"""Gallery-sharded bias measurement over the GPUs of one box.

One process per GPU.  Rank r owns a contiguous block of gallery rows (so "lowest global index"
-- the reference's tie rule -- does not depend on the shard layout), every rank holds all
queries.  The only exchange step of the path is small: per-query thresholds, label histograms
and the per-shard candidate lists travel through ``torch.distributed`` (NCCL all-reduce /
all-gather over NVLink); everything else is the C ABI's sharded stages
(``dvl_shard_topk`` -> ``dvl_shard_trim`` -> ``dvl_shard_pack`` -> ``dvl_merge_topk`` -> ``dvl_mean``).

    local streaming top-k  ->  all-reduce(max) tau        (no row below it can rank globally)
                               all-reduce(sum) label histogram, all-reduce(max) flags
    trim to tau            ->  all-reduce(max) list length
    pack                   ->  all-gather  [world, Q, len] u64 entries
    merge own query slice  ->  all-gather  per-query MaxSkew/NDKL
    mean (same reduction tree as one GPU: the result is bit-identical to the unsharded run)

A query whose boundary tie run does not fit a candidate list (say an all-zero query) is flagged
and re-ranked exactly: every rank computes that query's dense scores for its shard, the rows are
all-gathered and ranked by the full-sort route.

``evaluate_sharded`` takes a ``comm`` object with ``all_reduce_max/sum``, ``all_gather`` so that
the protocol can be exercised on CPU with gloo (tests) and emulated in one process (several
contexts on one GPU); ``TorchDistComm`` is the real thing.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _capi

__all__ = ["shard_bounds", "TorchDistComm", "SoloComm", "LocalComm", "evaluate_sharded", "evaluate_sharded_host", "evaluate_query_sharded",
           "measure_sharded", "speculative_list_len"]


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row blocks of ceil(n/world) rows (256-row aligned so that every shard but the
    last is a whole number of GEMM tiles)."""
    block = -(-n // world)
    block = -(-block // 256) * 256
    lo = min(n, rank * block)
    return lo, min(n, lo + block)


class TorchDistComm:
    """Collectives over a torch.distributed process group (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """[world, *t.shape]; every rank contributes the same shape."""
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        return out

    def all_gather_into(self, out: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Same into a caller-owned contiguous [world, *t.shape] buffer (no allocation on the way)."""
        self.dist.all_gather_into_tensor(out.view(-1), t.view(-1), group=self.group)
        return out


class SoloComm:
    """world = 1: the gallery-sharded protocol on ONE GPU holding the whole gallery (the N = 1 point of the
    strong-scaling curve).  The collectives degenerate to views; no process group is needed."""
    world = 1
    rank = 0

    @staticmethod
    def all_reduce_max(t: torch.Tensor) -> torch.Tensor:
        return t

    @staticmethod
    def all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
        return t

    @staticmethod
    def all_gather(t: torch.Tensor) -> torch.Tensor:
        return t.contiguous()[None]

    @staticmethod
    def all_gather_into(out: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        out.view(-1).copy_(t.view(-1))
        return out


class LocalComm:
    """The collectives as plain tensor ops over a list of per-rank tensors: used by
    ``evaluate_sharded_emulated`` to run all shards of a gallery from one process, stage by stage."""

    @staticmethod
    def reduce_max(ts: Sequence[torch.Tensor]) -> torch.Tensor:
        return torch.stack(list(ts)).max(dim=0).values

    @staticmethod
    def reduce_sum(ts: Sequence[torch.Tensor]) -> torch.Tensor:
        return torch.stack(list(ts)).sum(dim=0)

    @staticmethod
    def gather(ts: Sequence[torch.Tensor]) -> torch.Tensor:
        return torch.stack(list(ts)).contiguous()


def _flags_and_precision(opts: Optional[dict]):
    opts = dict(opts or {})
    precision = {"fp32": _capi.PRECISION_FP32, "bf16": _capi.PRECISION_BF16,
                 "fp32x3": _capi.PRECISION_FP32X3}[opts.get("precision", "fp32")]
    flags = 0
    if opts.get("normalize_gallery", False):
        flags |= _capi.FLAG_NORMALIZE_GALLERY
    if opts.get("normalize_queries", False):
        flags |= _capi.FLAG_NORMALIZE_QUERIES
    return precision, flags


def _query_slice(q: int, world: int, rank: int) -> Tuple[int, int]:
    block = -(-q // world)
    lo = min(q, rank * block)
    return lo, min(q, lo + block) - lo


# Packed list length learnt by a (world, Q, top_n, shard size) configuration whose speculative length
# was once too short (tie-heavy or label-sorted galleries): later calls start from it.
_learnt_list_len: Dict[tuple, int] = {}
MAX_STREAM_TOP_N = 8192      # dvl_shard_topk keeps lists of up to this many entries (csrc/api.cu plan_stream)
# exchange buffers of evaluate_query_sharded, per (context, device, queries per rank)
_query_shard_bufs: Dict[tuple, tuple] = {}


def speculative_list_len(top_n: int, rows_per_shard: int, n_global: int) -> int:
    """Rows of a shard expected among the global top_n (mu = top_n * shard / gallery) plus six
    binomial sigmas and a margin, rounded up to 8 entries.  Only a starting point: the true maximum
    travels with the results and a too-short guess is repaired by a second exchange."""
    mu = top_n * min(1.0, rows_per_shard / max(1, n_global))
    return int(-(-math.ceil(mu + 6.0 * math.sqrt(mu) + 32.0) // 8) * 8)


def _dense_rerank(ctx, comm, gallery_local: torch.Tensor, labels_local_u8: torch.Tensor, queries_sel: torch.Tensor,
                  n_global: int, n_classes: int, top_n: int, precision: int, flags: int, ranked_cap: int,
                  gallery_prepared: bool = False):
    """Exact ranking of a (small) set of queries over the whole sharded gallery by the full-sort route: every rank
    computes the dense scores of its shard, the score rows and the labels are all-gathered (contiguous row blocks,
    so concatenating the shards IS the gallery order) and ranked with ``dvl_rank_and_score``.  Used for the queries
    the streaming route flags (tie run longer than a candidate list) and for ``top_n`` beyond its 8192-entry lists;
    batched so that the gathered score matrix stays below ~2 GB."""
    world = comm.world
    dev = queries_sel.device
    n_local = gallery_local.shape[0]
    rows_blk = shard_bounds(n_global, world, 0)[1]
    lab_blk = torch.zeros((rows_blk,), dtype=torch.uint8, device=dev)
    lab_blk[:n_local] = labels_local_u8
    all_labels = comm.all_gather(lab_blk).reshape(-1)[:n_global].contiguous()
    f_total = queries_sel.shape[0]
    batch = max(1, min(f_total, int(2.0e9 // (4 * world * rows_blk))))
    out = {"per_query": torch.empty((f_total, 4), dtype=torch.float64, device=dev),
           "ranked_len": torch.empty((f_total,), dtype=torch.int32, device=dev),
           "ranked_idx": torch.full((f_total, ranked_cap), -1, dtype=torch.int32, device=dev) if ranked_cap else None}
    for b0 in range(0, f_total, batch):
        fq = queries_sel[b0:b0 + batch].contiguous()
        f = fq.shape[0]
        local_scores = torch.zeros((f, rows_blk), dtype=torch.float32, device=dev)
        if n_local:
            # right after shard_topk on the same, untouched shard its operand planes are still in the context
            reuse = _capi.FLAG_REUSE_GALLERY if (gallery_prepared and b0 == 0) else 0
            local_scores[:, :n_local] = ctx.similarity(gallery_local, fq, precision, flags | reuse)
        all_scores = comm.all_gather(local_scores)          # [world, F, rows_blk]
        full = all_scores.permute(1, 0, 2).reshape(f, -1)[:, :n_global].contiguous()
        rr = ctx.rank_and_score(full, all_labels, n_classes, top_n, ranked_cap)
        out["per_query"][b0:b0 + f] = rr["per_query"]
        out["ranked_len"][b0:b0 + f] = rr["ranked_len"]
        if ranked_cap:
            out["ranked_idx"][b0:b0 + f] = rr["ranked_idx"]
    return out


def evaluate_sharded_host(ctx, comm, gallery_local: torch.Tensor, queries: torch.Tensor, labels_local_u8: torch.Tensor,
                          index_base: int, n_global: int, n_classes: int, top_n: int, opts: Optional[dict] = None,
                          ranked_cap: int = 0) -> Dict[str, object]:
    """``evaluate_sharded`` for a shard that lives in HOST memory (pinned tensors: the embedding cache, or
    the encoders' host copy): the upload of the gallery rows is overlapped with the local streaming top-k
    (``dvl_shard_topk_host``); everything after it is the same exchange."""
    assert not gallery_local.is_cuda and not queries.is_cuda and not labels_local_u8.is_cuda
    return evaluate_sharded(ctx, comm, gallery_local, queries, labels_local_u8, index_base, n_global, n_classes, top_n,
                            opts, ranked_cap, host_inputs=True)


class _ExchangeBufs:
    """Persistent device buffers of one (context, Q, world, list length, ranked_cap) exchange: send / receive
    buffers of the three all-gathers, with typed views into this rank's result record
    ``[per_query f64 x block x 4][ranked_len i32 x block][flags u32 x block][need i32]`` that the merge
    kernels write directly.  Nothing is allocated, concatenated or sliced on the device inside a step."""

    def __init__(self, dev, q: int, world: int, length: int, ranked_cap: int):
        block = -(-q // world)
        words = 2 * q + 2 * (_capi.MAX_CLASSES + 1)
        self.block, self.length = block, length
        self.stats_send = torch.zeros((words,), dtype=torch.int32, device=dev)
        self.stats_all = torch.zeros((world, words), dtype=torch.int32, device=dev)
        self.packed = torch.full((q, length), -1, dtype=torch.int64, device=dev)
        self.gathered = torch.empty((world, q, length), dtype=torch.int64, device=dev)
        self.rec_bytes = -(-(40 * block + 4) // 16) * 16
        self.rec_send = torch.zeros((self.rec_bytes,), dtype=torch.uint8, device=dev)
        self.rec_all = torch.zeros((world, self.rec_bytes), dtype=torch.uint8, device=dev)
        self.pq = self.rec_send[:32 * block].view(torch.float64).view(block, 4)
        self.len = self.rec_send[32 * block:36 * block].view(torch.int32)
        self.flags = self.rec_send[36 * block:40 * block].view(torch.int32)
        self.need = self.rec_send[40 * block:40 * block + 4].view(torch.int32)
        self.idx_send = torch.full((block, ranked_cap), -1, dtype=torch.int32, device=dev) if ranked_cap else None
        self.idx_all = torch.empty((world, block, ranked_cap), dtype=torch.int32, device=dev) if ranked_cap else None

    def column(self, lo: int, hi: int, dtype, q: int):
        """Bytes [lo, hi) of every rank's record as one [world * block, ...] tensor of `dtype`, first q rows."""
        return self.rec_all[:, lo:hi].contiguous().view(dtype).reshape(-1)[:q] if dtype != torch.float64 else \
            self.rec_all[:, lo:hi].contiguous().view(dtype).reshape(-1, 4)[:q]


_exchange_bufs: Dict[tuple, _ExchangeBufs] = {}


def _bufs_for(ctx, dev, q: int, world: int, length: int, ranked_cap: int) -> _ExchangeBufs:
    key = (id(ctx), str(dev), q, world, length, ranked_cap)
    bufs = _exchange_bufs.get(key)
    if bufs is None:
        if len(_exchange_bufs) > 16:
            _exchange_bufs.clear()
        bufs = _exchange_bufs[key] = _ExchangeBufs(dev, q, world, length, ranked_cap)
    return bufs


def evaluate_sharded(ctx, comm, gallery_local: torch.Tensor, queries: torch.Tensor, labels_local_u8: torch.Tensor,
                     index_base: int, n_global: int, n_classes: int, top_n: int, opts: Optional[dict] = None,
                     ranked_cap: int = 0, host_inputs: bool = False) -> Dict[str, object]:
    """One rank's part of the sharded evaluation.  Returns the same dict on every rank:
    ``{"maxskew": {...}, "ndkl": {...}, "per_query": [Q,4] f64, "ranked_len": [Q] i32, "n_flagged": int}``.

    Three collectives and ONE host synchronisation on the usual path, all enqueued behind the local
    similarity GEMM, into persistent buffers (no allocation or tensor glue inside a step):

    1. all-gather of the per-shard statistics record (thresholds, overflow flags, label histogram),
       reduced by one kernel on every rank: tau = max over shards, histogram = sum;
    2. all-gather of the candidate lists, cut at tau and packed (one kernel) to a *speculative* common
       length: the rows of a shard expected among the global top_n plus six sigmas (224 entries instead of
       ~1000 at top-1000 on 8 GPUs -- the list gather is the one bandwidth-bound collective of the path).
       The merge verifies the guess per query: a list that was cut must end strictly behind the k-th entry
       of the union, else the query is marked and the exchange repeated with the untruncated length;
    3. all-gather of every rank's result record: the per-query results of its query slice, written there
       directly by the merge kernels, and the list length it really needed.
    ``dvl_shard_finish`` then reduces the records to the four means and a status word in one kernel and one
    device-to-host copy; only if a list was longer than the guess (or a tie run overflowed a list) is there
    a second, exact round."""
    precision, flags = _flags_and_precision(opts)
    world, rank = comm.world, comm.rank
    q = queries.shape[0]
    if top_n < 1:
        raise ZeroDivisionError("top_n must be >= 1")
    dev = torch.device("cuda", ctx.device) if host_inputs else queries.device
    if top_n > MAX_STREAM_TOP_N:
        # beyond the streaming top-k's shared-memory lists: every query is ranked exactly from gathered dense scores
        # (the reference takes any topn, measuring_bias.py:109-110; so does this, at full-sort cost)
        if host_inputs:
            gallery_local, queries, labels_local_u8 = gallery_local.to(dev), queries.to(dev), labels_local_u8.to(dev)
        rr = _dense_rerank(ctx, comm, gallery_local, labels_local_u8, queries, n_global, n_classes, top_n, precision,
                           flags, ranked_cap)
        mean = ctx.mean(rr["per_query"])
        out = {"maxskew": {"eq_opp": mean[0], "dem_par": mean[1]}, "ndkl": {"eq_opp": mean[2], "dem_par": mean[3]},
               "per_query": rr["per_query"], "ranked_len": rr["ranked_len"], "n_flagged": q, "list_len": 0}
        if ranked_cap:
            out["ranked_idx"] = rr["ranked_idx"]
        return out
    rows_per_shard = shard_bounds(n_global, world, 0)[1]
    key = (world, q, int(top_n), rows_per_shard, n_global)
    length = _learnt_list_len.get(key) or speculative_list_len(top_n, rows_per_shard, n_global)
    q0, qn = _query_slice(q, world, rank)
    have_rows = gallery_local.shape[0] > 0       # a rank may own no rows when the gallery is tiny
    bufs = _bufs_for(ctx, dev, q, world, length, ranked_cap)

    # 1. local streaming top-k -> this rank's statistics record -> reduced over the ranks
    if have_rows:
        local_topk = ctx.shard_topk_host if host_inputs else ctx.shard_topk
        local_topk(gallery_local, queries, labels_local_u8, index_base, n_classes, top_n, precision, flags, outputs=False)
        ctx.shard_stats(bufs.stats_send)
    else:
        bufs.stats_send.zero_()
        bufs.stats_send[:q] = torch.full((q,), float("-inf"), dtype=torch.float32, device=dev).view(torch.int32)
    comm.all_gather_into(bufs.stats_all, bufs.stats_send)
    ctx.shard_reduce(bufs.stats_all, q)

    def exchange_and_merge(bufs: _ExchangeBufs):
        # 2. what can still rank globally, packed to bufs.length entries per query
        if have_rows:
            ctx.shard_trim_pack(bufs.packed, bufs.need)
        else:
            bufs.packed.fill_(-1)
            bufs.need.zero_()
        comm.all_gather_into(bufs.gathered, bufs.packed)           # [world, Q, length]
        # 3. merge this rank's slice of the queries straight into its result record
        if ranked_cap:
            bufs.idx_send.fill_(-1)
        if qn > 0:
            ctx.merge_topk_into(bufs.gathered, q0, qn, n_global, n_classes, top_n, bufs.pq[:qn], bufs.len[:qn],
                                bufs.flags[:qn], bufs.idx_send[:qn] if ranked_cap else None)
        comm.all_gather_into(bufs.rec_all, bufs.rec_send)          # [world, rec_bytes]
        if ranked_cap:
            comm.all_gather_into(bufs.idx_all, bufs.idx_send)
        # the caller's copies of the per-query results (the buffers are reused by the next call) are enqueued
        # BEFORE the synchronisation: the host prepares them while the device is still busy, instead of after
        # it has gone idle
        block = bufs.block
        copies = (bufs.column(0, 32 * block, torch.float64, q).clone(),
                  bufs.column(32 * block, 36 * block, torch.int32, q).clone(),
                  bufs.idx_all.reshape(-1, ranked_cap)[:q].clone() if ranked_cap else None)
        return ctx.shard_finish(bufs.rec_all, bufs.block, q), copies      # the path's host synchronisation

    (mean, (any_flag, bad_label, need_max)), copies = exchange_and_merge(bufs)
    if bad_label:
        raise _capi.DvlError(-1, f"a label is >= n_classes ({n_classes})")
    if need_max > length:
        # a shard holds more of the global top_n than the guess allowed for (every rank sees the same
        # need_max, so every rank takes this branch): exchange again with the exact length and remember it
        length = int(-(-need_max // 8) * 8)
        _learnt_list_len[key] = max(length, _learnt_list_len.get(key, 0))
        stats_all = bufs.stats_all
        bufs = _bufs_for(ctx, dev, q, world, length, ranked_cap)
        bufs.stats_all.copy_(stats_all)
        (mean, (any_flag, bad_label, need_max)), copies = exchange_and_merge(bufs)
    block = bufs.block
    per_query, ranked_len, ranked_idx = copies
    n_flagged = 0
    # exact re-rank of flagged queries (tie run longer than a candidate list); rare
    if any_flag:
        oflags = bufs.stats_all[:, q:2 * q].max(dim=0).values
        merge_flags = bufs.column(36 * block, 40 * block, torch.int32, q)
        flag_mask = (oflags != 0) | ((merge_flags & 1) != 0)
        if host_inputs:      # rare path: the dense scores of the flagged queries need the shard on the device
            gallery_local, queries, labels_local_u8 = gallery_local.to(dev), queries.to(dev), labels_local_u8.to(dev)
        flagged = torch.nonzero(flag_mask).reshape(-1)
        n_flagged = int(flagged.numel())
        rr = _dense_rerank(ctx, comm, gallery_local, labels_local_u8, queries[flagged].contiguous(), n_global, n_classes,
                           top_n, precision, flags, ranked_cap, gallery_prepared=have_rows and not host_inputs)
        per_query[flagged] = rr["per_query"]
        ranked_len[flagged] = rr["ranked_len"]
        if ranked_cap:
            ranked_idx[flagged] = rr["ranked_idx"]
        mean = ctx.mean(per_query)
    out = {"maxskew": {"eq_opp": mean[0], "dem_par": mean[1]}, "ndkl": {"eq_opp": mean[2], "dem_par": mean[3]},
           "per_query": per_query, "ranked_len": ranked_len, "n_flagged": n_flagged, "list_len": length}
    if ranked_cap:
        out["ranked_idx"] = ranked_idx
    return out


def evaluate_sharded_emulated(contexts: Sequence, gallery: torch.Tensor, queries: torch.Tensor, labels_u8: torch.Tensor,
                              n_classes: int, top_n: int, opts: Optional[dict] = None, ranked_cap: int = 0,
                              lean: bool = True, length: Optional[int] = None):
    """All `len(contexts)` ranks in ONE process on one GPU, stage by stage (no kernel ever waits on
    another rank): the same C-ABI stage calls as `evaluate_sharded` (``lean``: the statistics record /
    reduce / trim+pack / merge-into-record / finish sequence it uses; otherwise the separate
    trim / pack / merge / mean stages), with the collectives replaced by tensor ops.  Used by the GPU tests."""
    if not lean:
        return _evaluate_sharded_emulated_stages(contexts, gallery, queries, labels_u8, n_classes, top_n, opts, ranked_cap)
    precision, flags = _flags_and_precision(opts)
    world = len(contexts)
    n, q = gallery.shape[0], queries.shape[0]
    dev = queries.device
    bounds = [shard_bounds(n, world, r) for r in range(world)]
    shards = [(gallery[lo:hi].contiguous(), labels_u8[lo:hi].contiguous(), lo) for lo, hi in bounds]
    live = [r for r in range(world) if shards[r][0].shape[0] > 0]
    if length is None:
        length = speculative_list_len(top_n, bounds[0][1], n)
    bufs = [_ExchangeBufs(dev, q, world, length, ranked_cap) for _ in range(world)]
    for r in range(world):
        if r in live:
            g, l, lo = shards[r]
            contexts[r].shard_topk(g, queries, l, lo, n_classes, top_n, precision, flags)
            contexts[r].shard_stats(bufs[r].stats_send)
        else:
            bufs[r].stats_send[:q] = torch.full((q,), float("-inf"), dtype=torch.float32, device=dev).view(torch.int32)
    stats_all = torch.stack([b.stats_send for b in bufs]).contiguous()
    for r in range(world):
        bufs[r].stats_all.copy_(stats_all)
        contexts[r].shard_reduce(bufs[r].stats_all, q)

    def exchange(bufs):
        for r in live:
            contexts[r].shard_trim_pack(bufs[r].packed, bufs[r].need)
        gathered = torch.stack([b.packed for b in bufs]).contiguous()
        for r in range(world):
            q0, qn = _query_slice(q, world, r)
            if qn > 0:
                contexts[r].merge_topk_into(gathered, q0, qn, n, n_classes, top_n, bufs[r].pq[:qn], bufs[r].len[:qn],
                                            bufs[r].flags[:qn], bufs[r].idx_send[:qn] if ranked_cap else None)
        rec_all = torch.stack([b.rec_send for b in bufs]).contiguous()
        bufs[0].rec_all.copy_(rec_all)
        if ranked_cap:
            bufs[0].idx_all.copy_(torch.stack([b.idx_send for b in bufs]))
        return contexts[0].shard_finish(bufs[0].rec_all, bufs[0].block, q)

    mean, (any_flag, bad_label, need_max) = exchange(bufs)
    if bad_label:
        raise _capi.DvlError(-1, f"a label is >= n_classes ({n_classes})")
    if need_max > length:
        length = int(-(-need_max // 8) * 8)
        bufs = [_ExchangeBufs(dev, q, world, length, ranked_cap) for _ in range(world)]
        bufs[0].stats_all.copy_(stats_all)
        mean, (any_flag, bad_label, need_max) = exchange(bufs)
    block = bufs[0].block
    per_query = bufs[0].column(0, 32 * block, torch.float64, q).clone()
    ranked_len = bufs[0].column(32 * block, 36 * block, torch.int32, q).clone()
    ranked_idx = bufs[0].idx_all.reshape(-1, ranked_cap)[:q].clone() if ranked_cap else None
    n_flagged = 0
    if any_flag:
        oflags = stats_all[:, q:2 * q].max(dim=0).values
        flag_mask = (oflags != 0) | ((bufs[0].column(36 * block, 40 * block, torch.int32, q) & 1) != 0)
        flagged = torch.nonzero(flag_mask).reshape(-1)
        n_flagged = int(flagged.numel())
        fq = queries[flagged].contiguous()
        parts = [contexts[r].similarity(shards[r][0], fq, precision, flags) for r in live]
        rr = contexts[0].rank_and_score(torch.cat(parts, dim=1).contiguous(), labels_u8, n_classes, top_n, ranked_cap)
        per_query[flagged] = rr["per_query"]
        ranked_len[flagged] = rr["ranked_len"]
        if ranked_cap:
            ranked_idx[flagged] = rr["ranked_idx"]
        mean = contexts[0].mean(per_query)
    out = {"maxskew": {"eq_opp": mean[0], "dem_par": mean[1]}, "ndkl": {"eq_opp": mean[2], "dem_par": mean[3]},
           "per_query": per_query, "ranked_len": ranked_len, "n_flagged": n_flagged, "list_len": length}
    if ranked_cap:
        out["ranked_idx"] = ranked_idx
    return out


def _evaluate_sharded_emulated_stages(contexts: Sequence, gallery: torch.Tensor, queries: torch.Tensor,
                                      labels_u8: torch.Tensor, n_classes: int, top_n: int, opts: Optional[dict] = None,
                                      ranked_cap: int = 0):
    """All `len(contexts)` ranks in ONE process on one GPU, stage by stage (no kernel ever waits on
    another rank): the same C-ABI stage calls as `evaluate_sharded`, with the collectives replaced
    by tensor ops.  Used by the GPU tests."""
    precision, flags = _flags_and_precision(opts)
    world = len(contexts)
    n, q = gallery.shape[0], queries.shape[0]
    bounds = [shard_bounds(n, world, r) for r in range(world)]
    shards = [(gallery[lo:hi].contiguous(), labels_u8[lo:hi].contiguous(), lo) for lo, hi in bounds]
    s1 = [ctx.shard_topk(g, queries, l, lo, n_classes, top_n, precision, flags) if g.shape[0] else None
          for ctx, (g, l, lo) in zip(contexts, shards)]
    live = [r for r in range(world) if s1[r] is not None]
    tau = LocalComm.reduce_max([s1[r][0] for r in live])
    oflags = LocalComm.reduce_max([s1[r][1] for r in live])
    counts = LocalComm.reduce_sum([s1[r][2] for r in live])
    if int(counts[_capi.MAX_CLASSES].item()) != 0:
        raise _capi.DvlError(-1, f"a label is >= n_classes ({n_classes})")
    cnts = {r: contexts[r].shard_trim(tau) for r in live}
    length = max(1, max(int(c.max().item()) for c in cnts.values()))
    empty = torch.full((q, length), -1, dtype=torch.int64, device=queries.device)
    gathered = LocalComm.gather([contexts[r].shard_pack(cnts[r], length) if r in cnts else empty for r in range(world)])
    per_query = torch.zeros((q, 4), dtype=torch.float64, device=queries.device)
    ranked_len = torch.zeros((q,), dtype=torch.int32, device=queries.device)
    merge_flags = torch.zeros((q,), dtype=torch.int32, device=queries.device)
    ranked_idx = torch.full((q, ranked_cap), -1, dtype=torch.int32, device=queries.device) if ranked_cap else None
    for r in range(world):
        q0, qn = _query_slice(q, world, r)
        if qn <= 0:
            continue
        res = contexts[r].merge_topk(gathered, q0, qn, n, counts[:_capi.MAX_CLASSES].contiguous(), n_classes, top_n,
                                     ranked_cap)
        per_query[q0:q0 + qn] = res["per_query"]
        ranked_len[q0:q0 + qn] = res["ranked_len"]
        merge_flags[q0:q0 + qn] = res["flags"]
        if ranked_cap:
            ranked_idx[q0:q0 + qn] = res["ranked_idx"]
    flagged = torch.nonzero((oflags != 0) | ((merge_flags & 1) != 0)).reshape(-1)   # bit 1 (list cut short) cannot be set: exact lengths
    if flagged.numel():
        fq = queries[flagged].contiguous()
        parts = [contexts[r].similarity(shards[r][0], fq, precision, flags) for r in live]
        full = torch.cat(parts, dim=1).contiguous()
        rr = contexts[0].rank_and_score(full, labels_u8, n_classes, top_n, ranked_cap)
        per_query[flagged] = rr["per_query"]
        ranked_len[flagged] = rr["ranked_len"]
        if ranked_cap:
            ranked_idx[flagged] = rr["ranked_idx"]
    mean = contexts[0].mean(per_query)
    out = {"maxskew": {"eq_opp": mean[0], "dem_par": mean[1]}, "ndkl": {"eq_opp": mean[2], "dem_par": mean[3]},
           "per_query": per_query, "ranked_len": ranked_len, "n_flagged": int(flagged.numel()), "list_len": length}
    if ranked_cap:
        out["ranked_idx"] = ranked_idx
    return out


def evaluate_query_sharded(ctx, comm, gallery: torch.Tensor, queries: torch.Tensor, labels_u8: torch.Tensor,
                           n_classes: int, top_n: int, opts: Optional[dict] = None, ranked_cap: int = 0
                           ) -> Dict[str, object]:
    """The ``top_n >= N`` regime (the ``measure_bias`` default: the whole gallery is ranked) on several
    GPUs: queries are independent and the gallery is small (FairFace: 55-222 MB), so every rank holds
    the whole gallery and ranks a contiguous slice of the queries by the full-sort route.  No data-path
    collective: one all-gather of the per-query results at the end, then the same fixed-tree mean as on
    one GPU, so the numbers are bit-identical to the unsharded run.  Also the route for ``top_n`` beyond
    the streaming top-k's 8192-entry lists."""
    precision, flags = _flags_and_precision(opts)
    world, rank = comm.world, comm.rank
    q = queries.shape[0]
    q0, qn = _query_slice(q, world, rank)
    block = -(-q // world)
    dev = queries.device
    # Persistent exchange buffers per (context, shape).  The kernels write the per-query results straight
    # into the send buffer, and a caller that repeats the evaluation hands the C ABI the same pointers
    # every time (the CUDA-graph replay of the full-sort route keys on them).  The step after the local
    # ranking is pure launch latency (a few tiny kernels + one NCCL call on an idle GPU), so it is kept
    # to: all-gather into a persistent [world, block, 4] buffer -- which IS the per-query matrix -- and
    # the mean.  The ranked lengths travel in a second all-gather only when asked for
    # (opts["ranked_len"], default on) and the per-query matrix is copied out of the reused receive buffer
    # only when asked for (opts["per_query"], default on); measure_bias itself only needs the means.
    want_len = bool((opts or {}).get("ranked_len", True)) or ranked_cap > 0
    key = (id(ctx), str(dev), block, world)
    bufs = _query_shard_bufs.get(key)
    if bufs is None:
        bufs = (torch.zeros((block, 4), dtype=torch.float64, device=dev),
                torch.zeros((block,), dtype=torch.int32, device=dev),
                torch.zeros((world, block, 4), dtype=torch.float64, device=dev))
        _query_shard_bufs[key] = bufs
    send, len_buf, recv = bufs
    idx_blk = torch.full((block, ranked_cap), -1, dtype=torch.int32, device=dev) if ranked_cap else None
    if qn > 0:
        res = ctx.eval_ranking(gallery, queries[q0:q0 + qn], labels_u8, n_classes, top_n, precision,
                               flags | _capi.FLAG_FULL_SORT, ranked_cap=ranked_cap,
                               out_per_query=send[:qn], out_ranked_len=len_buf[:qn])
        if ranked_cap:
            idx_blk[:qn] = res["ranked_idx"]
    comm.all_gather_into(recv, send)                       # [world, block, 4]
    per_query = recv.view(world * block, 4)[:q]            # contiguous: a prefix of the receive buffer
    mean = ctx.mean(per_query)
    out = {"maxskew": {"eq_opp": mean[0], "dem_par": mean[1]}, "ndkl": {"eq_opp": mean[2], "dem_par": mean[3]},
           "n_flagged": 0}
    if bool((opts or {}).get("per_query", True)):
        out["per_query"] = per_query.clone()               # the receive buffer is reused by the next call
    if want_len:
        out["ranked_len"] = comm.all_gather(len_buf).reshape(-1)[:q].contiguous()
    if ranked_cap:
        out["ranked_idx"] = comm.all_gather(idx_blk).reshape(-1, ranked_cap)[:q].contiguous()
    return out


def measure_sharded(labels_local, gallery_local: torch.Tensor, queries: torch.Tensor, index_base: int, n_global: int,
                    topn: Union[int, float] = 1000, opts: Optional[dict] = None, group=None):
    """Convenience wrapper under torchrun: labels (ints, local rows) -> the measure_bias dict."""
    lab = np.asarray(labels_local)
    if lab.size and (int(lab.min()) < 0 or int(lab.max()) >= _capi.MAX_CLASSES):
        raise ValueError(f"labels must be class codes 0..{_capi.MAX_CLASSES - 1} (datasets.gen_labels); "
                         "re-encode them before sharding")
    # a rank may own no rows (tiny gallery): its class count is whatever the other ranks see
    n_classes = int(opts.get("n_classes")) if opts and "n_classes" in opts else (int(lab.max()) + 1 if lab.size else 1)
    top_n = math.ceil(n_global * topn) if isinstance(topn, float) else int(topn)
    dev = gallery_local.device
    comm = TorchDistComm(group)
    nc = comm.all_reduce_max(torch.tensor([n_classes], dtype=torch.int64, device=dev))
    n_classes = int(nc.item())
    ctx = _capi.context(dev.index)
    res = evaluate_sharded(ctx, comm, gallery_local.contiguous(), queries.contiguous(),
                           torch.from_numpy(lab.astype(np.uint8)).to(dev), index_base, n_global, n_classes, top_n, opts)
    return {"maxskew": res["maxskew"], "ndkl": res["ndkl"]}
