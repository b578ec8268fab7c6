"""Row-by-row CPU port of the reference's hot path.  TEST / BASELINE INFRASTRUCTURE ONLY.

Where ``oracle/spec.py`` is the vectorised checker, this file restates the reference the way
the reference runs: one query at a time, a pandas ``DataFrame{"score","label"}`` per query,
``nlargest(keep="all")`` for the ranking, and a Python loop with two tiny numpy KL evaluations
per ranked row -- i.e. the same cost structure as ``/root/reference/debias_clip/
measuring_bias.py`` (:113-117 per query, :32-37 per ranked row).  ``bench.py`` times it as the
CPU baseline (``cpu_baseline.kind == "port"``, and the ``--impl reference`` arm) because the
reference itself is not present on the GPU box.  Checked against the golden vectors in
``tests/test_oracle.py``.
"""
from __future__ import annotations

import math
from collections import Counter
from typing import Dict, List, Union

import numpy as np
import pandas as pd
import torch


def _kl(p: np.ndarray, q: np.ndarray) -> float:
    # measuring_bias.py:18-19 -- zero-probability classes contribute nothing
    return np.sum(np.where(p != 0, p * (np.log(p) - np.log(q)), 0))


def ndkl_one_query(frame: pd.DataFrame, top_n: int) -> Dict[str, float]:
    """measuring_bias.py:17-44."""
    per_label = sorted(Counter(frame.label).items())
    counts = [c for _, c in per_label]
    n_labels = len(counts)
    target = {
        "eq_opp": np.array([1 / n_labels] * n_labels),
        "dem_par": np.array([max(c, 1) / len(frame) for c in counts]),
    }
    ranked = frame.nlargest(top_n, columns="score", keep="all")
    seen = np.zeros(n_labels)
    total = {name: 0.0 for name in target}
    position = 0
    with np.errstate(divide="ignore", invalid="ignore"):      # log(0) of classes not seen yet
        for _, row in ranked.iterrows():
            position += 1
            seen[int(row["label"])] += 1
            for name, dist in target.items():
                total[name] += _kl(seen / position, dist) / math.log2(position + 1)
    z = sum(1 / math.log2(i + 1) for i in range(1, top_n + 1))
    return {name: float(value / z) for name, value in total.items()}


def maxskew_one_query(frame: pd.DataFrame, top_n: int) -> Dict[str, float]:
    """measuring_bias.py:47-72."""
    overall = Counter(frame.label)
    ranked = frame.nlargest(top_n, columns="score", keep="all")
    in_top = Counter(ranked.label)
    worst = {"eq_opp": 0, "dem_par": 0}
    for label, count in overall.items():
        share = in_top[label] / top_n
        if share == 0:
            share = 1 / top_n
        for name, desired in (("eq_opp", 1 / len(overall)), ("dem_par", count / len(frame))):
            worst[name] = max(worst[name], math.log(share) - math.log(desired))
    return worst


def eval_ranking(labels_list, image_embeddings: torch.Tensor, prompts_embeddings: torch.Tensor,
                 evaluation: str = "maxskew", topn: Union[int, float] = 1.0) -> Dict[str, float]:
    """measuring_bias.py:102-119."""
    assert evaluation in ("maxskew", "ndkl")
    fn = maxskew_one_query if evaluation == "maxskew" else ndkl_one_query
    if isinstance(topn, float):
        topn = math.ceil(len(image_embeddings) * topn)
    collected: Dict[str, List[float]] = {}
    for query in prompts_embeddings:
        scores = (image_embeddings.float() @ query.float()).cpu().numpy().flatten()
        frame = pd.DataFrame({"score": scores, "label": labels_list})
        for name, value in fn(frame, topn).items():
            collected.setdefault(name, []).append(value)
    return {name: sum(v) / len(v) for name, v in collected.items()}


def measure(labels_list, image_embeddings, prompts_embeddings, topn: Union[int, float] = 1.0):
    """Both evaluations, as measure_bias runs them (measuring_bias.py:152-156)."""
    return {ev: eval_ranking(labels_list, image_embeddings, prompts_embeddings, ev, topn)
            for ev in ("maxskew", "ndkl")}
