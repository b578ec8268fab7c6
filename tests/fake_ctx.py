"""A CPU stand-in for the sharded C-ABI stages, built on the oracle, so that the collective
protocol of debias_vision_lang_b200.sharded can be exercised with gloo on CPU (world_size 2).
Test infrastructure only."""
import math

import numpy as np
import torch

from oracle import spec

MAX_CLASSES = 16


def _key(s: float) -> int:
    u = int(np.float32(s).view(np.uint32))
    if s != s:
        return 0xFFFFFFFF
    if s == 0.0:
        u = 0
    return u if (u & 0x80000000) else (u ^ 0x7FFFFFFF)


def _entry(s, gidx, label):
    return (_key(s) << 32) | (int(gidx) << 4) | int(label)


def _to_i64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


class FakeCtx:
    device = 0

    def __init__(self, list_cap=64):
        self.list_cap = list_cap
        self.lists = None

    def shard_topk(self, gallery, queries, labels_u8, index_base, n_classes, top_n, precision=0, flags=0, outputs=True):
        q = queries.shape[0]
        lab = labels_u8.numpy()
        tau = torch.full((q,), float("-inf"), dtype=torch.float32)
        oflags = torch.zeros((q,), dtype=torch.int32)
        self.lists = []
        for i in range(q):
            s = spec.similarity_f32(gallery, queries[i])
            order = spec.rank_indices(s, top_n)
            if len(order) > self.list_cap:
                oflags[i] = 1
                tau[i] = float("inf")
                self.lists.append([])
                continue
            self.lists.append([(float(s[j]), _entry(s[j], index_base + j, lab[j])) for j in order])
            if len(s) >= top_n and len(order) >= top_n:
                tau[i] = float(s[order[top_n - 1]])
        counts = torch.zeros((MAX_CLASSES + 1,), dtype=torch.int64)
        counts[:n_classes] = torch.from_numpy(np.bincount(lab, minlength=n_classes)[:n_classes].astype(np.int64))
        counts[MAX_CLASSES] = int((lab >= n_classes).any())
        self._tau, self._oflags, self._counts = tau, oflags, counts
        return tau, oflags, counts, self.list_cap

    # ---- the lean exchange (mirrors _capi.Context) ----
    def shard_stats(self, out_stats):
        q = len(self.lists)
        out_stats[:q] = self._tau.view(torch.int32)
        out_stats[q:2 * q] = self._oflags
        out_stats[2 * q:] = self._counts.view(torch.int32)

    def shard_reduce(self, stats_all, q):
        self._tau_bound = stats_all[:, :q].contiguous().view(torch.float32).max(dim=0).values
        self._flags_any = stats_all[:, q:2 * q].max(dim=0).values
        self._counts_global = stats_all[:, 2 * q:].contiguous().view(torch.int64).sum(dim=0)

    def shard_trim_pack(self, out_packed, out_need):
        cnt = self.shard_trim(self._tau_bound)
        out_packed.copy_(self.shard_pack(cnt, out_packed.shape[1]))
        out_need[0] = int(cnt.max())

    def merge_topk_into(self, gathered, q0, qn, n_global, n_classes, top_n, per_query, ranked_len, flags, ranked_idx=None):
        cap = ranked_idx.shape[1] if ranked_idx is not None else 0
        res = self.merge_topk(gathered, q0, qn, n_global, self._counts_global[:MAX_CLASSES], n_classes, top_n, cap)
        per_query.copy_(res["per_query"])
        ranked_len.copy_(res["ranked_len"])
        flags.copy_(res["flags"])
        if cap:
            ranked_idx.copy_(res["ranked_idx"])

    def shard_finish(self, records_all, block, q):
        world = records_all.shape[0]
        pq = records_all[:, :32 * block].contiguous().view(torch.float64).reshape(-1, 4)[:q]
        mflags = records_all[:, 36 * block:40 * block].contiguous().view(torch.int32).reshape(-1)[:q]
        need = records_all[:, 40 * block:40 * block + 4].contiguous().view(torch.int32).reshape(-1)
        any_flag = int(bool((mflags != 0).any()) or bool((self._flags_any != 0).any()))
        return self.mean(pq), (any_flag, int(self._counts_global[MAX_CLASSES] != 0), int(need.max()))

    def shard_trim(self, tau_bound):
        return torch.tensor([sum(1 for s, _ in lst if s >= float(tau_bound[i])) for i, lst in enumerate(self.lists)],
                            dtype=torch.int32)

    def shard_pack(self, cnt, length):
        out = torch.full((len(self.lists), length), -1, dtype=torch.int64)
        for i, lst in enumerate(self.lists):
            for j in range(min(int(cnt[i]), length)):      # a too-short speculative length truncates
                out[i, j] = _to_i64(lst[j][1])
        return out

    def merge_topk(self, gathered, q0, qn, n_global, class_counts, n_classes, top_n, ranked_cap=0):
        per_query = torch.zeros((qn, 4), dtype=torch.float64)
        ranked_len = torch.zeros((qn,), dtype=torch.int32)
        oflags = torch.zeros((qn,), dtype=torch.int32)
        ranked_idx = torch.full((qn, ranked_cap), -1, dtype=torch.int32) if ranked_cap else None
        counts = class_counts.numpy()[:n_classes]
        values = np.arange(n_classes)
        for i in range(qn):
            ent = sorted(int(v) & ((1 << 64) - 1) for v in gathered[:, q0 + i, :].reshape(-1).tolist() if int(v) != -1)
            ent = [e for e in ent if (e >> 32) != 0xFFFFFFFF]
            if len(ent) > top_n:
                kth = ent[top_n - 1] >> 32
                ent = [e for e in ent if (e >> 32) <= kth]
            if len(ent) > self.list_cap * 4:
                oflags[i] = 1
                continue
            ranked = np.array([e & 0xF for e in ent], dtype=np.int64)
            ms = spec.maxskew_from_ranked(ranked, values, counts, n_global, top_n)
            nd = spec.ndkl_from_ranked(ranked, counts, n_global, top_n)
            per_query[i] = torch.tensor([ms["eq_opp"], ms["dem_par"], nd["eq_opp"], nd["dem_par"]], dtype=torch.float64)
            ranked_len[i] = len(ent)
            if ranked_cap:
                idx = [(e >> 4) & 0x0FFFFFFF for e in ent][:ranked_cap]
                ranked_idx[i, :len(idx)] = torch.tensor(idx, dtype=torch.int32)
        return dict(per_query=per_query, ranked_len=ranked_len, ranked_idx=ranked_idx, flags=oflags)

    def similarity(self, gallery, queries, precision=0, flags=0):
        return torch.stack([torch.from_numpy(spec.similarity_f32(gallery, queries[i])) for i in range(queries.shape[0])])

    def rank_and_score(self, scores, labels_u8, n_classes, top_n, ranked_cap=0):
        pq, orders = spec.eval_scores_matrix(scores.numpy(), labels_u8.numpy().astype(np.int64), top_n)
        ranked_idx = None
        if ranked_cap:
            ranked_idx = torch.full((len(orders), ranked_cap), -1, dtype=torch.int32)
            for i, o in enumerate(orders):
                ranked_idx[i, :min(len(o), ranked_cap)] = torch.from_numpy(o[:ranked_cap].astype(np.int32))
        return dict(per_query=torch.from_numpy(pq), ranked_len=torch.tensor([len(o) for o in orders], dtype=torch.int32),
                    ranked_idx=ranked_idx, mean=pq.mean(axis=0).tolist())

    def eval_ranking(self, gallery, queries, labels_u8, n_classes, top_n, precision=0, flags=0, want_per_query=False,
                     ranked_cap=0, want_scores=False, out_per_query=None, out_ranked_len=None):
        scores = self.similarity(gallery, queries)
        res = self.rank_and_score(scores, labels_u8, n_classes, top_n, ranked_cap)
        if out_per_query is not None:          # caller-owned output buffers, as in _capi.Context.eval_ranking
            out_per_query.view(-1, 4).copy_(res["per_query"])
        if out_ranked_len is not None:
            out_ranked_len.copy_(res["ranked_len"])
        return res

    def mean(self, per_query):
        return [float(x) for x in (per_query.sum(dim=0) / per_query.shape[0]).tolist()]
