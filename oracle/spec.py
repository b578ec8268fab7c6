"""CPU oracle for the bias-measurement hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``debias_vision_lang_b200``) never does.

It is a vectorised numpy/float64 restatement of the reference algorithm
(``/root/reference/debias_clip/measuring_bias.py``) and of the pandas ``nlargest``
semantics that algorithm relies on (pandas 3.0.2 ``core/methods/selectn.py:106-190``,
an un-vendored, unpinned dependency of the reference: ``requirements.txt:8``).

Parity pin: the reference has no tests and no golden vectors of its own
(SURVEY.md section 4), so the pin is made from outputs of the reference itself,
executed in the build container through ``oracle/ref_shim.py`` and committed as
``tests/golden/*.json`` by ``oracle/make_golden.py``.  ``tests/test_oracle.py`` checks this
file against every one of those vectors, and (when ``/root/reference`` is mounted)
against the live reference on random tie-heavy cases.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

__all__ = [
    "resolve_topn", "similarity_f32", "rank_indices", "class_counts", "maxskew_from_ranked",
    "ndkl_from_ranked", "ndkl_normaliser", "eval_query", "eval_ranking", "measure",
    "split_bf16x3", "split_f16x2", "similarity_bf16", "similarity_split", "similarity_split_f16x2",
]


# --------------------------------------------------------------------------------------
# top-n resolution                                       measuring_bias.py:107-110
# --------------------------------------------------------------------------------------
def resolve_topn(topn: Union[int, float], n_gallery: int) -> int:
    """A float is a proportion of the gallery (ceil), an int is taken as is."""
    if isinstance(topn, float):
        return int(math.ceil(n_gallery * topn))
    return int(topn)


# --------------------------------------------------------------------------------------
# similarity                                             measuring_bias.py:114
# --------------------------------------------------------------------------------------
def similarity_f32(gallery, query) -> np.ndarray:
    """fp32 mat-vec of the raw gallery against ONE query row, exactly the call the
    reference makes (torch CPU ``float() @ float()``), so that on one machine the
    scores are bit-identical to the reference's.  Falls back to numpy sgemv when
    torch is not importable."""
    try:
        import torch
        g = torch.as_tensor(np.asarray(gallery)) if not hasattr(gallery, "float") else gallery
        t = torch.as_tensor(np.asarray(query)) if not hasattr(query, "float") else query
        t = t.reshape(-1)  # a 1-D query: `q.T` of the reference is the identity on it
        return (g.float() @ t.float()).cpu().numpy().flatten()
    except ImportError:  # pragma: no cover
        return (np.asarray(gallery, np.float32) @ np.asarray(query, np.float32)).astype(np.float32)


# --------------------------------------------------------------------------------------
# ranking = DataFrame.nlargest(top_n, "score", keep="all")
#                                    measuring_bias.py:29,57 -> selectn.py:106-190
# --------------------------------------------------------------------------------------
def rank_indices(scores: np.ndarray, top_n: int) -> np.ndarray:
    """Row indices in ranked order: score descending, lowest index first among equal
    scores (-0.0 == +0.0), NaN rows dropped (selectn.py:133-134), and *every* row
    tied with the top_n-th score kept (keep="all", selectn.py:176-182) so the result
    can be longer than top_n.  top_n >= len returns the full stable descending order
    (selectn.py:124-130)."""
    s = np.asarray(scores)
    valid = np.flatnonzero(~np.isnan(s))
    if top_n <= 0 or valid.size == 0:
        return valid[:0]
    # stable argsort of the negated values == (score desc, index asc); -0.0/+0.0 compare equal
    order = valid[np.argsort(-s[valid], kind="stable")]
    if top_n >= order.size:
        return order
    kth = s[order[top_n - 1]]
    m = top_n + int(np.count_nonzero(s[order[top_n:]] == kth))
    return order[:m]


def class_counts(labels: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Counter(df.label) sorted by label (measuring_bias.py:23,56): the distinct label
    values present and how often each occurs."""
    values, counts = np.unique(np.asarray(labels), return_counts=True)
    return values, counts.astype(np.int64)


# --------------------------------------------------------------------------------------
# MaxSkew@k                                              measuring_bias.py:47-72
# --------------------------------------------------------------------------------------
def maxskew_from_ranked(ranked_labels: np.ndarray, values: np.ndarray, counts: np.ndarray,
                        n_gallery: int, top_n: int) -> Dict[str, float]:
    """``ranked_labels`` are the labels of the m >= top_n ranked rows.  p_c = h_c/top_n
    (can exceed 1 when boundary ties were kept), zero -> 1/top_n (:64-65), skew
    ln p - ln d against eq_opp d=1/C (C = #distinct labels present) and dem_par
    d=n_c/N; the running max starts at 0 (:54,69)."""
    out = {"eq_opp": 0.0, "dem_par": 0.0}
    n_classes = len(values)
    rl = np.asarray(ranked_labels)
    for value, n_c in zip(values, counts):
        h = int(np.count_nonzero(rl == value))
        p = h / top_n
        if p == 0:
            p = 1 / top_n
        lp = math.log(p)
        out["eq_opp"] = max(out["eq_opp"], lp - math.log(1 / n_classes))
        out["dem_par"] = max(out["dem_par"], lp - math.log(int(n_c) / n_gallery))
    return out


# --------------------------------------------------------------------------------------
# NDKL                                                   measuring_bias.py:17-44
# --------------------------------------------------------------------------------------
def ndkl_normaliser(top_n: int) -> float:
    """Z = sum_{i=1..top_n} 1/log2(i+1)  (:39), summed sequentially like the reference."""
    i = np.arange(1, top_n + 1, dtype=np.float64)
    return float(np.cumsum(1.0 / np.log2(i + 1.0))[-1]) if top_n > 0 else 0.0


def ndkl_from_ranked(ranked_labels: np.ndarray, counts: np.ndarray, n_gallery: int,
                     top_n: int) -> Dict[str, float]:
    """Labels must already be the contiguous ints 0..C-1 the reference indexes its count
    array with (:33-34).  For every prefix i of the m ranked rows: KL(h(i)/i || d) over
    the classes with non-zero prefix count (:18-19), discounted by log2(i+1) (:37),
    summed over *all m* rows but normalised by Z(top_n) (:39-42)."""
    rl = np.asarray(ranked_labels, dtype=np.int64)
    n_classes = len(counts)
    if rl.size and (rl.max() >= n_classes or rl.min() < 0):
        raise IndexError("label outside 0..C-1: the reference indexes top_n_label_counts[label]")
    dists = {
        "eq_opp": np.full(n_classes, 1.0 / n_classes, dtype=np.float64),
        "dem_par": np.maximum(counts, 1).astype(np.float64) / float(n_gallery),
    }
    m = rl.size
    z = ndkl_normaliser(top_n)
    if m == 0:
        return {k: (0.0 / z if z else float("nan")) for k in dists}
    out = {}
    idx = np.arange(1, m + 1, dtype=np.float64)
    discount = np.log2(idx + 1.0)
    # prefix histograms, processed in slabs so that m = 1e7 does not need m*C float64 at once
    slab = max(1, min(m, (1 << 22) // max(n_classes, 1)))
    acc = {k: 0.0 for k in dists}
    carry = np.zeros(n_classes, dtype=np.int64)
    for lo in range(0, m, slab):
        hi = min(m, lo + slab)
        onehot = (rl[lo:hi, None] == np.arange(n_classes)[None, :])
        h = np.cumsum(onehot, axis=0, dtype=np.int64) + carry[None, :]
        carry = h[-1].copy()
        p = h.astype(np.float64) / idx[lo:hi, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            logp = np.log(p)
            for name, d in dists.items():
                term = np.where(p != 0, p * (logp - np.log(d)[None, :]), 0.0)
                kl = term.sum(axis=1)
                acc[name] += float(np.sum(kl / discount[lo:hi]))
    for name in dists:
        out[name] = acc[name] / z
    return out


# --------------------------------------------------------------------------------------
# per-query evaluation and the mean over queries         measuring_bias.py:102-119
# --------------------------------------------------------------------------------------
def eval_query(scores: np.ndarray, labels: np.ndarray, evaluation: str, top_n: int) -> Dict[str, float]:
    values, counts = class_counts(labels)
    order = rank_indices(scores, top_n)
    ranked = np.asarray(labels)[order]
    if evaluation == "maxskew":
        return maxskew_from_ranked(ranked, values, counts, len(labels), top_n)
    if evaluation == "ndkl":
        return ndkl_from_ranked(ranked, counts, len(labels), top_n)
    raise AssertionError(evaluation)


def eval_ranking(labels_list, image_embeddings, prompts_embeddings, evaluation: str = "maxskew",
                 topn: Union[int, float] = 1.0, similarity=similarity_f32,
                 return_per_query: bool = False):
    """Restatement of ``eval_ranking`` (measuring_bias.py:102-119): per query, fp32
    similarities of the *un-normalised* gallery (only the query side was normalised
    upstream, :79), rank, metric; sequential mean over queries (:119)."""
    assert evaluation in ("maxskew", "ndkl")
    labels = np.asarray(labels_list)
    n = len(image_embeddings)
    top_n = resolve_topn(topn, n)
    per_query = {"eq_opp": [], "dem_par": []}
    for qi in range(len(prompts_embeddings)):
        s = similarity(image_embeddings, prompts_embeddings[qi])
        r = eval_query(s, labels, evaluation, top_n)
        for k in per_query:
            per_query[k].append(r[k])
    mean = {k: sum(v) / len(v) for k, v in per_query.items()}
    if return_per_query:
        return mean, {k: np.asarray(v, dtype=np.float64) for k, v in per_query.items()}
    return mean


def measure(labels_list, image_embeddings, prompts_embeddings, topn: Union[int, float] = 1.0,
            similarity=similarity_f32) -> Dict[str, Dict[str, float]]:
    """Both evaluations, in the dict shape ``measure_bias`` returns (measuring_bias.py:152-156)."""
    return {ev: eval_ranking(labels_list, image_embeddings, prompts_embeddings, ev, topn, similarity)
            for ev in ("maxskew", "ndkl")}


def eval_scores_matrix(scores_qn: np.ndarray, labels: np.ndarray, top_n: int):
    """Stage-isolated oracle: rank + both metrics for an explicit [Q, N] fp32 score matrix
    (used to check the CUDA ranking/metric stages on the CUDA path's own similarities).
    Returns (per_query float64 [Q,4] in the order maxskew eq_opp, maxskew dem_par,
    ndkl eq_opp, ndkl dem_par; list of ranked index arrays)."""
    labels = np.asarray(labels)
    values, counts = class_counts(labels)
    out = np.zeros((scores_qn.shape[0], 4), dtype=np.float64)
    orders = []
    for qi in range(scores_qn.shape[0]):
        order = rank_indices(scores_qn[qi], top_n)
        ranked = labels[order]
        ms = maxskew_from_ranked(ranked, values, counts, len(labels), top_n)
        nd = ndkl_from_ranked(ranked, counts, len(labels), top_n)
        out[qi] = (ms["eq_opp"], ms["dem_par"], nd["eq_opp"], nd["dem_par"])
        orders.append(order)
    return out, orders


# --------------------------------------------------------------------------------------
# models of the CUDA path's arithmetic (for tolerance tests, not part of the reference)
# --------------------------------------------------------------------------------------
def _bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bf16, returned as fp32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(x)
    out[nan] = np.float32("nan")
    return out


def split_bf16x3(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """x == b1 + b2 + b3 exactly (each bf16-representable): the operand split the CUDA
    fp32 mode feeds the bf16 tensor cores with."""
    x = np.asarray(x, dtype=np.float32)
    b1 = _bf16_round(x)
    r1 = (x - b1).astype(np.float32)
    b2 = _bf16_round(r1)
    r2 = (r1 - b2).astype(np.float32)
    b3 = _bf16_round(r2)
    return b1, b2, b3


def split_f16x2(x: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The default fp32 mode's operand split: every row is scaled by the power of two that brings its
    largest magnitude into [2^14, 2^15), then x * 2^e = h0 + h1 + eps with h0, h1 fp16 and
    |eps| <= 2^-22 |x * 2^e| (round to nearest twice).  Returns (h0, h1, down) with down = 2^-e per row;
    all-zero and non-finite rows keep scale 1."""
    x = np.asarray(x, dtype=np.float32)
    amax = np.abs(x).max(axis=-1)
    ok = np.isfinite(amax) & (amax > 0)
    be = ((np.where(ok, amax, 1.0).astype(np.float32).view(np.uint32) >> 23) & 0xFF).astype(np.int64)
    e = np.clip(14 - (be - 127), -126, 126)
    e = np.where(ok, e, 0)
    up = np.ldexp(np.float32(1.0), e).astype(np.float32)
    down = np.ldexp(np.float32(1.0), -e).astype(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        xs = (x * up[..., None]).astype(np.float32)
        h0 = xs.astype(np.float16)
        h1 = (xs - h0.astype(np.float32)).astype(np.float16)
    return h0, h1, down


def similarity_split_f16x2(gallery: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """[Q, N] float64 value of the three product terms the default fp32 mode keeps (h1*h0', h0*h1',
    h0*h0'), descaled; differs from the exact product by at most ~3 * 2^-22 of |g||t|."""
    g0, g1, gd = (a.astype(np.float64) for a in split_f16x2(gallery))
    t0, t1, td = (a.astype(np.float64) for a in split_f16x2(queries))
    acc = t1 @ g0.T + t0 @ g1.T + t0 @ g0.T
    return acc * td[:, None] * gd[None, :]


def similarity_bf16(gallery: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """[Q, N] similarities with both operands rounded to bf16 and products accumulated in
    float64 (the exact value the bf16 tensor-core mode approximates)."""
    g = _bf16_round(np.asarray(gallery, np.float32)).astype(np.float64)
    t = _bf16_round(np.asarray(queries, np.float32)).astype(np.float64)
    return t @ g.T


def similarity_split(gallery: np.ndarray, queries: np.ndarray) -> np.ndarray:
    """[Q, N] float64 value of the six bf16x3 product terms the fp32 mode keeps
    (i + j <= 4); differs from the exact product by O(2^-24)."""
    gs = [p.astype(np.float64) for p in split_bf16x3(gallery)]
    ts = [p.astype(np.float64) for p in split_bf16x3(queries)]
    acc = 0.0
    for i, j in ((0, 2), (1, 1), (2, 0), (0, 1), (1, 0), (0, 0)):
        acc = acc + ts[i] @ gs[j].T
    return acc
