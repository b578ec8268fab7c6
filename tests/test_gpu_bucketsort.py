"""The full-sort route's fast path (csrc/bucketsort.cu: two levels of score-linear bucketing) against the
oracle on score distributions chosen to stress it: heavy tails (outliers stretch the first-level range),
clusters with empty gaps, massive exact ties (bucket overflow -> the call is repeated with the radix sort),
constant scores, NaN / +-inf / -0.0 / denormal scores, tiny and ragged segment lengths, top_n < n (boundary
ties need the sorted keys) and ranked-index export.  Stage-isolated: dvl_rank_and_score on explicit score
matrices, compared with oracle/spec.py's ranking + metrics of the same scores (bit-exact index lists)."""
import zlib

import numpy as np
import pytest
import torch

from debias_vision_lang_b200 import _capi, synthetic
from oracle import spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    c = _capi.Context(0)
    yield c
    c.close()


def _scores(kind, q, n, rng):
    if kind == "normal":
        s = rng.standard_normal((q, n)) * 0.4 + 1.0
    elif kind == "cauchy":                       # outliers: the sampled range is far wider than the bulk
        s = rng.standard_cauchy((q, n))
    elif kind == "exponential":
        s = -rng.exponential(1.0, (q, n))
    elif kind == "bimodal_gap":
        s = np.where(rng.rand(q, n) < 0.5, rng.standard_normal((q, n)) * 1e-3 - 50.0,
                     rng.standard_normal((q, n)) * 1e-3 + 50.0)
    elif kind == "small_ints":                   # 7 distinct values: every bucket overflows -> radix fallback
        s = rng.randint(-3, 4, (q, n)).astype(np.float64) / 8.0
    elif kind == "constant":
        s = np.full((q, n), 0.25)
    elif kind == "around_zero":                  # scores straddling zero incl. denormals and signed zeros
        s = rng.standard_normal((q, n)) * 1e-3
        s[:, ::17] = 0.0
        s[:, 5::17] = -0.0
        s[:, 3::29] = 1e-42
        s[:, 4::29] = -1e-42
    elif kind == "specials":
        s = rng.standard_normal((q, n))
        s[:, ::11] = np.nan
        s[:, 1::13] = np.inf
        s[:, 2::13] = -np.inf
        s[:, 3::13] = 3.0e38
        s[:, 4::13] = -3.0e38
        if q > 1:
            s[1, :] = np.nan                     # a query without a single finite score
    elif kind == "mostly_ties":                  # half the rows share one value, the rest are spread out
        s = rng.standard_normal((q, n))
        s[:, ::2] = 0.5
    else:
        raise ValueError(kind)
    return s.astype(np.float32)


@pytest.mark.parametrize("kind", ["normal", "cauchy", "exponential", "bimodal_gap", "small_ints", "constant",
                                  "around_zero", "specials", "mostly_ties"])
@pytest.mark.parametrize("n,q,c,top", [(30_011, 5, 7, None), (30_011, 3, 2, 1000), (4_097, 4, 3, None)])
def test_bucket_sort_route_matches_oracle(ctx, kind, n, q, c, top):
    rng = np.random.RandomState(zlib.crc32(f"{kind}:{n}:{top}".encode()))
    scores = _scores(kind, q, n, rng)
    labels = synthetic.make_labels(n, c, 3)
    top_n = n if top is None else top
    want_pq, want_order = spec.eval_scores_matrix(scores, labels, top_n)
    res = ctx.rank_and_score(torch.from_numpy(scores).cuda(), torch.from_numpy(labels.astype(np.uint8)).cuda(), c, top_n,
                             ranked_cap=n)
    lens = res["ranked_len"].cpu().numpy()
    idx = res["ranked_idx"].cpu().numpy()
    for qi in range(q):
        assert lens[qi] == len(want_order[qi]), (kind, qi, lens[qi], len(want_order[qi]))
        assert np.array_equal(idx[qi, :lens[qi]], want_order[qi]), (kind, qi)
    got = res["per_query"].cpu().numpy()
    ok = np.isfinite(want_pq)
    assert np.abs(got[ok] - want_pq[ok]).max() <= 1e-9 if ok.any() else True


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 2048, 4096, 8191, 100_003])
def test_bucket_sort_segment_lengths(ctx, n):
    rng = np.random.RandomState(n)
    q, c = 3, 2
    scores = (rng.standard_normal((q, n)) * 0.3).astype(np.float32)
    labels = synthetic.make_labels(max(n, c), c, 5)[:n]
    if n >= c:
        labels[:c] = np.arange(c)
    want_pq, want_order = spec.eval_scores_matrix(scores, labels, n)
    res = ctx.rank_and_score(torch.from_numpy(scores).cuda(), torch.from_numpy(labels.astype(np.uint8)).cuda(), c, n,
                             ranked_cap=n)
    idx = res["ranked_idx"].cpu().numpy()
    for qi in range(q):
        assert np.array_equal(idx[qi], want_order[qi]), (n, qi)
    assert np.abs(res["per_query"].cpu().numpy() - want_pq).max() <= 1e-9


def test_bucket_sort_equals_radix_sort_on_config2_shape(ctx):
    """BASELINE config 2's shape end to end (108,501 x 512 x 40 queries, C = 7, top_n = N): the default route
    (bucket sort) and the radix sort (DVL_FLAG_RADIX_SORT) give identical ranked
    index lists and per-query results equal to 1e-12, in direct launches and in the CUDA-graph replays of either."""
    w = synthetic.baseline_workload("cfg2", device="cuda", q_override=40)
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    n = w.gallery.shape[0]
    runs = {}
    for name, flags in (("bucket", 0), ("radix", _capi.FLAG_RADIX_SORT)):
        outs = [ctx.eval_ranking(w.gallery, w.queries, lab, 7, n, _capi.PRECISION_FP32, flags, want_per_query=True,
                                 ranked_cap=n) for _ in range(4)]          # direct, capture, replay, replay
        for o in outs[1:]:
            assert torch.equal(o["per_query"], outs[0]["per_query"]) and torch.equal(o["ranked_idx"], outs[0]["ranked_idx"])
        runs[name] = outs[0]
    assert torch.equal(runs["bucket"]["ranked_idx"], runs["radix"]["ranked_idx"])
    assert torch.equal(runs["bucket"]["ranked_len"], runs["radix"]["ranked_len"])
    # (1e-12, not bit equality: with DVL_FUSE_METRICS=1 the bucket route sums the NDKL terms per bucket instead of
    # per 1024-row slice -- same fp64 terms, different grouping)
    assert (runs["bucket"]["per_query"] - runs["radix"]["per_query"]).abs().max().item() <= 1e-12
    assert max(abs(a - b) for a, b in zip(runs["bucket"]["mean"], runs["radix"]["mean"])) <= 1e-12
    want = spec.measure(w.labels, w.gallery.cpu()[:, :], w.queries.cpu()[:3], 1.0)    # three queries through the oracle
    got3 = runs["bucket"]["per_query"][:3].cpu().numpy().mean(axis=0)
    assert abs(got3[2] - want["ndkl"]["eq_opp"]) <= 1e-6 and abs(got3[3] - want["ndkl"]["dem_par"]) <= 1e-6


def test_bucket_overflow_falls_back_inside_graph_replays(ctx):
    """Tie-heavy scores overflow a bucket: every call (direct, captured, replayed) must notice and return the
    radix sort's result."""
    w = synthetic.adversarial(20_000, 64, 12, 2, seed=9)      # query 0 is all zeros, tiny score alphabet
    g, t = w.gallery.cuda(), w.queries.cuda()
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    want = spec.measure(w.labels, w.gallery, w.queries, 1.0)
    for _ in range(5):
        m = ctx.eval_ranking(g, t, lab, 2, 20_000, _capi.PRECISION_FP32)["mean"]
        assert abs(m[2] - want["ndkl"]["eq_opp"]) <= 1e-9 and abs(m[3] - want["ndkl"]["dem_par"]) <= 1e-9
        assert abs(m[0] - want["maxskew"]["eq_opp"]) <= 1e-9


def test_query_batching_of_the_full_sort_route_matches_one_batch(tmp_path):
    """The full-sort route ranks queries in batches bounded by a workspace budget (DVL_WS_BUDGET_MB, read once per
    process): a run squeezed into ~10 queries per batch -- bucket sort and radix sort, with ranked indices and
    scores on request -- must equal the unbatched run bit for bit."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    script = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
from debias_vision_lang_b200 import _capi, synthetic
w = synthetic.clip_like(30_011, 64, 47, 7, seed=5, device="cuda")
lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
ctx = _capi.Context(0)
out = {}
for name, flags, topn in (("bucket", 0, 30_011), ("radix", _capi.FLAG_RADIX_SORT, 30_011), ("cut", _capi.FLAG_FULL_SORT, 700)):
    r = ctx.eval_ranking(w.gallery, w.queries, lab, 7, topn, _capi.PRECISION_FP32, flags, want_per_query=True,
                         ranked_cap=1500, want_scores=(name == "bucket"))
    out[name] = dict(mean=r["mean"], pq=r["per_query"].cpu().numpy().tobytes().hex(),
                     idx=r["ranked_idx"].cpu().numpy().tobytes().hex(), len=r["ranked_len"].cpu().tolist(),
                     scores=(r["scores"].double().sum().item() if r["scores"] is not None else None))
print("RESULT" + json.dumps(out))
''' % ROOT
    results = []
    for budget in (None, "8"):
        env = dict(os.environ)
        env.pop("DVL_WS_BUDGET_MB", None)
        if budget:
            env["DVL_WS_BUDGET_MB"] = budget
        p = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        results.append(json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("RESULT")][0][6:]))
    assert results[0] == results[1]


def test_more_segments_than_one_grid_dimension_holds(ctx):
    """70,000 queries (more than the 65,535 a grid's y dimension takes: the launches are split): every query is one
    of seven score rows, so its result must equal that row's -- which is checked against the oracle."""
    n, q, c = 300, 70_000, 3
    rng = np.random.RandomState(12)
    base = (rng.standard_normal((7, n)) * 0.5).astype(np.float32)
    base[3, ::4] = 0.25                                   # some exact ties
    which = rng.randint(0, 7, size=q)
    which[:7] = np.arange(7)
    scores = torch.from_numpy(base).cuda()[torch.from_numpy(which).cuda()].contiguous()
    labels = synthetic.make_labels(n, c, 8)
    lab = torch.from_numpy(labels.astype(np.uint8)).cuda()
    for top_n in (n, 40):
        want, _ = spec.eval_scores_matrix(base, labels, top_n)
        res = ctx.rank_and_score(scores, lab, c, top_n)
        got = res["per_query"]
        assert np.abs(got[:7].cpu().numpy() - want).max() <= 1e-12
        assert torch.equal(got, got[:7][torch.from_numpy(which).cuda()])
