"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the
golden vectors the reference produced.  Run on the B200 box with `-m gpu`.

Bars (BASELINE.json north_star): rankings bit-exact (lowest-index tie-break, boundary ties
kept); similarities within 1e-5 (fp32 mode) / 2e-2 (bf16 mode) norm-wise relative
(|ds| <= tol*|g||t|: a purely relative bound is meaningless near s = 0);
MaxSkew and NDKL within 1e-6 absolute (in practice ~1e-15)."""
import math

import numpy as np
import pytest
import torch

from conftest import build_ranking_workload, load_golden, u32_to_f32
from debias_vision_lang_b200 import _capi, measuring_bias, synthetic
import debias_vision_lang_b200 as dvl
from oracle import spec

pytestmark = pytest.mark.gpu

METRIC_TOL = 1e-6      # the bar
TIGHT = 1e-12          # what float64 on both sides actually gives on identical rankings


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available()
    return _capi.context(0)


def _flat(d):
    return [d["maxskew"]["eq_opp"], d["maxskew"]["dem_par"], d["ndkl"]["eq_opp"], d["ndkl"]["dem_par"]]


def test_stage_golden_vectors_rank_and_metrics(ctx, stage_cases):
    """Every (scores, labels, top_n) case the reference was run on: ranked order bit-exact,
    four metrics to 1e-12 -- ties, signed zeros, NaN, +-inf, denormals, top_n > n included."""
    for case in stage_cases:
        s = u32_to_f32(case["scores_u32"])
        lab = np.asarray(case["labels"])
        codes = np.searchsorted(np.unique(lab), lab).astype(np.uint8)
        c = len(np.unique(lab))
        cap = max(1, len(case["order"]))
        res = ctx.rank_and_score(torch.from_numpy(s.copy()).cuda()[None, :], torch.from_numpy(codes).cuda(), c,
                                 case["top_n"], ranked_cap=cap)
        m = int(res["ranked_len"][0])
        assert m == len(case["order"]), case["name"]
        assert res["ranked_idx"][0, :m].cpu().tolist() == case["order"], case["name"]
        got = res["per_query"][0].cpu().tolist()
        # the reference's ndkl needs labels 0..C-1; golden ndkl values exist only for such cases
        want = [case["maxskew_eq_opp"], case["maxskew_dem_par"], case["ndkl_eq_opp"], case["ndkl_dem_par"]]
        for g, w_ in zip(got, want):
            if math.isnan(w_):
                continue
            assert abs(g - w_) <= TIGHT * max(1.0, abs(w_)), (case["name"], got, want)


def test_ranking_golden_vectors_end_to_end(ranking_cases):
    """eval_ranking-level vectors of the reference on exactly-summable (dyadic) embeddings, through the
    public drop-in API, fp32 mode: similarity + ranking + metrics + mean."""
    for case in ranking_cases:
        w, topn = build_ranking_workload(case)
        g, t = w.gallery.cuda(), w.queries.cuda()
        for ev in ("maxskew", "ndkl"):
            got = dvl.eval_ranking(w.labels, g, t, ev, topn)
            for k in ("eq_opp", "dem_par"):
                assert abs(got[k] - case["expected"][ev][k]) <= TIGHT, (case, ev, k, got)


@pytest.mark.parametrize("n,d,q", [(256, 64, 128), (1000, 128, 130), (5000, 512, 319), (3001, 768, 257),
                                   (777, 200, 33), (70000, 1024, 64)])
def test_similarity_tolerances(ctx, n, d, q):
    torch.manual_seed(n + d + q)
    g = torch.randn(n, d, device="cuda") * (1.0 + 9.0 * torch.rand(n, 1, device="cuda"))
    t = torch.randn(q, d, device="cuda")
    t = t / t.norm(dim=-1, keepdim=True)
    exact = t.double() @ g.double().T
    scale = g.double().norm(dim=-1)[None, :] * t.double().norm(dim=-1)[:, None]
    s32 = ctx.similarity(g, t, _capi.PRECISION_FP32)
    assert ((s32.double() - exact).abs() / scale).max().item() <= 1e-5
    # in practice the scaled fp16x2 split (bound 3 * 2^-22 of |g||t|) is as accurate as an fp32 dot product
    assert ((s32.double() - exact).abs() / scale).max().item() <= 1e-6
    # and the exact bf16x3 split is tighter still
    sx3 = ctx.similarity(g, t, _capi.PRECISION_FP32X3)
    assert ((sx3.double() - exact).abs() / scale).max().item() <= 5e-7
    sbf = ctx.similarity(g, t, _capi.PRECISION_BF16)
    assert ((sbf.double() - exact).abs() / scale).max().item() <= 2e-2
    # and bf16 mode equals the product of the bf16-rounded operands up to fp32 accumulation
    model = t.bfloat16().double() @ g.bfloat16().double().T
    assert ((sbf.double() - model).abs() / scale).max().item() <= 2e-6


def test_similarity_exact_on_dyadic_grid(ctx):
    w = synthetic.dyadic_grid(5000, 512, 70, 2, seed=12)
    g, t = w.gallery.cuda(), w.queries.cuda()
    exact = (t.double() @ g.double().T)
    # |k| <= 127 fits one bf16 / fp16 plane exactly, and the row scales of the fp16 split are powers of two
    for prec in (_capi.PRECISION_FP32, _capi.PRECISION_BF16, _capi.PRECISION_FP32X3):
        s = ctx.similarity(g, t, prec)
        assert torch.equal(s.double(), exact)


@pytest.mark.parametrize("prec,tol", [(_capi.PRECISION_FP32, 1e-6), (_capi.PRECISION_FP32X3, 5e-7)])
def test_similarity_over_the_fp32_range(ctx, prec, tol):
    """Row magnitudes from 1e-30 to 1e+30 (far outside what fp16 holds), elements spanning five decades
    inside a row, an all-zero row, D not a multiple of 4 (scalar-load path) and D > 512 (two-pass path):
    the per-row power-of-two scale keeps the fp32 mode's error at that of an fp32 dot product."""
    for n, d, q in ((2048, 300, 40), (1024, 1026, 24), (512, 777, 8)):
        torch.manual_seed(d)
        g = torch.randn(n, d, device="cuda") * torch.logspace(-30, 6, n, device="cuda")[:, None]
        g = g * torch.logspace(0, -5, d, device="cuda")[None, :]
        g[7] = 0.0
        t = torch.randn(q, d, device="cuda") * torch.logspace(-6, 20, q, device="cuda")[:, None]
        exact = t.double() @ g.double().T
        scale = g.double().norm(dim=-1)[None, :] * t.double().norm(dim=-1)[:, None]
        s = ctx.similarity(g, t, prec)
        err = (s.double() - exact).abs() / scale.clamp_min(1e-300)
        err[:, 7] = (s[:, 7] != 0).double()          # the zero row scores exactly zero
        assert torch.isfinite(s).all()
        assert err.max().item() <= tol, (n, d, q, err.max().item())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_similarity_of_half_precision_inputs(ctx, dtype):
    """CLIP runs fp16 on CUDA: 16-bit embeddings go through the same split kernels (vector
    loads when D % 4 == 0, element loads otherwise, mixed gallery/query dtypes) and every precision mode
    must reproduce the fp32 product of the 16-bit values."""
    for n, d, q in ((3000, 512, 150), (1500, 130, 20)):
        torch.manual_seed(n)
        g = (torch.randn(n, d, device="cuda") * 3).to(dtype)
        t = torch.randn(q, d, device="cuda")
        t = (t / t.norm(dim=-1, keepdim=True)).to(dtype)
        exact = t.double() @ g.double().T
        scale = g.double().norm(dim=-1)[None, :] * t.double().norm(dim=-1)[:, None]
        for prec, tol in ((_capi.PRECISION_FP32, 1e-6), (_capi.PRECISION_FP32X3, 5e-7), (_capi.PRECISION_BF16, 2e-2)):
            s = ctx.similarity(g, t, prec)
            assert ((s.double() - exact).abs() / scale).max().item() <= tol, (dtype, n, d, q, prec)
        # fp32 queries against a 16-bit gallery, as measure_bias hands them over (queries are .float()ed, :81)
        s = ctx.similarity(g, t.float(), _capi.PRECISION_FP32)
        assert ((s.double() - exact).abs() / scale).max().item() <= 1e-6


def test_fused_normalisation_flags(ctx):
    torch.manual_seed(5)
    g = torch.randn(3000, 256, device="cuda") * 7
    t = torch.randn(40, 256, device="cuda") * 3
    s = ctx.similarity(g, t, _capi.PRECISION_FP32, _capi.FLAG_NORMALIZE_QUERIES)
    want = (t / t.norm(dim=-1, keepdim=True)).double() @ g.double().T
    assert ((s.double() - want).abs() / g.double().norm(dim=-1)[None, :]).max().item() <= 1e-5
    s = ctx.similarity(g, t, _capi.PRECISION_FP32, _capi.FLAG_NORMALIZE_QUERIES | _capi.FLAG_NORMALIZE_GALLERY)
    want = (t / t.norm(dim=-1, keepdim=True)).double() @ (g / g.norm(dim=-1, keepdim=True)).double().T
    assert (s.double() - want).abs().max().item() <= 1e-5


@pytest.mark.parametrize("n,q,c,topn", [(10954, 16, 2, 1000), (10954, 8, 2, 10954), (30011, 6, 7, 30011),
                                        (30011, 6, 9, 500), (4096, 5, 3, 1), (513, 3, 2, 600),
                                        (20483, 4, 16, 20483), (9001, 3, 12, 1027), (1025, 2, 8, 1025),
                                        (127, 2, 5, 127), (2049, 2, 13, 130)])
def test_stage_isolation_random_scores(ctx, n, q, c, topn):
    """The CUDA similarities (random, un-tied data) fed back through the oracle's ranking+metrics must
    give the identical ranked index lists and metrics."""
    w = synthetic.clip_like(n, 128, q, c, seed=n + q)
    g, t = w.gallery.cuda(), w.queries.cuda()
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    res = ctx.eval_ranking(g, t, lab, c, topn, _capi.PRECISION_FP32, want_per_query=True,
                           ranked_cap=min(n, topn) + 64, want_scores=True)
    scores = res["scores"].cpu().numpy()
    want, orders = spec.eval_scores_matrix(scores, w.labels, topn)
    lens = res["ranked_len"].cpu().numpy()
    idx = res["ranked_idx"].cpu().numpy()
    for i in range(q):
        assert lens[i] == len(orders[i])
        assert np.array_equal(idx[i, :lens[i]], orders[i])
    assert np.abs(res["per_query"].cpu().numpy() - want).max() <= TIGHT
    assert np.abs(np.asarray(res["mean"]) - want.mean(axis=0)).max() <= TIGHT


def test_metrics_match_oracle_on_clip_like_data():
    """End to end against the oracle's own fp32 similarities (CPU BLAS order): metrics within the 1e-6
    bar even though individual near-tied ranks may legitimately differ."""
    w = synthetic.clip_like(20000, 512, 48, 7, seed=9)
    for topn in (1000, 1.0):
        want = spec.measure(w.labels, w.gallery, w.queries, topn)
        got = dvl.evaluate(w.labels, w.gallery.cuda(), w.queries.cuda(), topn)
        assert max(abs(a - b) for a, b in zip(_flat(got), _flat(want))) <= METRIC_TOL


def test_bf16_mode_matches_oracle_on_bf16_rounded_inputs():
    w = synthetic.clip_like(12000, 256, 32, 2, seed=10)
    g16 = w.gallery.bfloat16().float()
    t16 = w.queries.bfloat16().float()
    want = spec.measure(w.labels, g16, t16, 500)
    got = dvl.evaluate(w.labels, w.gallery.cuda(), w.queries.cuda(), 500, {"precision": "bf16"})
    assert max(abs(a - b) for a, b in zip(_flat(got), _flat(want))) <= METRIC_TOL
    # bf16 gallery handed in directly gives the same numbers
    got2 = dvl.evaluate(w.labels, w.gallery.cuda().bfloat16(), w.queries.cuda(), 500, {"precision": "bf16"})
    assert _flat(got2) == _flat(got)


def test_gv1_regression_vector():
    gv = load_golden("gv1.json")
    torch.manual_seed(gv["seed"])
    img = torch.randn(gv["n"], gv["d"])
    txt = torch.randn(gv["q"], gv["d"])
    txt /= txt.norm(dim=-1, keepdim=True)
    labels = np.random.RandomState(gv["seed"]).randint(0, gv["c"], size=gv["n"])
    g, t = img.cuda(), txt.cuda()
    for key, want in gv["expected"].items():
        ev, topn = key.split("@")
        topn = float(topn) if "." in topn else int(topn)
        got = dvl.eval_ranking(labels, g, t, ev, topn)
        for k in want:
            assert abs(got[k] - want[k]) <= METRIC_TOL, (key, k, got[k], want[k])


def test_dataframe_entry_points_match_golden(stage_cases):
    import pandas as pd
    for case in stage_cases[:40]:
        lab = np.asarray(case["labels"])
        if not np.array_equal(np.unique(lab), np.arange(len(np.unique(lab)))):
            continue
        df = pd.DataFrame({"score": u32_to_f32(case["scores_u32"]), "label": lab})
        ms = dvl.compute_skew_metrics(df, case["top_n"])
        nd = dvl.normalized_discounted_KL(df, case["top_n"])
        assert abs(ms["maxskew_eq_opp"] - case["maxskew_eq_opp"]) <= TIGHT
        assert abs(ms["maxskew_dem_par"] - case["maxskew_dem_par"]) <= TIGHT
        if not math.isnan(case["ndkl_eq_opp"]):
            assert abs(nd["ndkl_eq_opp"] - case["ndkl_eq_opp"]) <= TIGHT
            assert abs(nd["ndkl_dem_par"] - case["ndkl_dem_par"]) <= TIGHT


def test_dataframe_entry_points_rank_float64_scores_exactly():
    """float64 scores that differ below float32 resolution are NOT ties (the reference's nlargest compares the
    column as it is): the drop-in ranks them through an order- and tie-preserving surrogate."""
    import pandas as pd
    rng = np.random.RandomState(3)
    n, top_n = 400, 50
    base = rng.randint(0, 40, size=n).astype(np.float64) / 8.0           # heavy float32-level ties ...
    scores = base + rng.randint(0, 3, size=n) * 1e-12                      # ... split (or not) far below fp32 resolution
    labels = synthetic.make_labels(n, 3, 4)
    df = pd.DataFrame({"score": scores, "label": labels})
    ms = dvl.compute_skew_metrics(df, top_n)
    nd = dvl.normalized_discounted_KL(df, top_n)
    want_ms = spec.eval_query(scores, labels, "maxskew", top_n)
    want_nd = spec.eval_query(scores, labels, "ndkl", top_n)
    assert abs(ms["maxskew_eq_opp"] - want_ms["eq_opp"]) <= TIGHT and abs(ms["maxskew_dem_par"] - want_ms["dem_par"]) <= TIGHT
    assert abs(nd["ndkl_eq_opp"] - want_nd["eq_opp"]) <= TIGHT and abs(nd["ndkl_dem_par"] - want_nd["dem_par"]) <= TIGHT
    # and the float32 view of the same column would have been a different ranking
    as32 = spec.eval_query(scores.astype(np.float32), labels, "ndkl", top_n)
    assert abs(as32["eq_opp"] - want_nd["eq_opp"]) > 1e-6


def test_non_finite_embedding_elements_behave_like_an_fp32_product(ctx):
    """A gallery row with a +-inf element scores +-inf (by the sign of the query's component; NaN where it is 0 or
    where +inf and -inf meet), a NaN element scores NaN -- what the reference's fp32 mat-vec gives
    (measuring_bias.py:114) -- in every precision mode, and such rows rank first / last / are dropped."""
    torch.manual_seed(9)
    n, d, q = 600, 96, 5
    g = torch.randn(n, d, device="cuda")
    t = torch.randn(q, d, device="cuda")
    t[:, 7] = torch.tensor([1.5, -0.5, 0.0, 2.0, -3.0], device="cuda")
    t[:, 11] = torch.tensor([1.0, 1.0, 1.0, -1.0, 0.25], device="cuda")
    g[5, 7] = float("inf")
    g[6, 7] = float("-inf")
    g[9, 7] = float("inf"); g[9, 11] = float("-inf")
    g[12, 3] = float("nan")
    want = t @ g.t()                                   # fp32 reference semantics for the special rows
    rows = [5, 6, 9, 12]
    for prec in (_capi.PRECISION_FP32, _capi.PRECISION_FP32X3, _capi.PRECISION_BF16):
        got = ctx.similarity(g, t, prec)
        for r in rows:
            w, h = want[:, r], got[:, r]
            assert torch.equal(torch.isnan(w), torch.isnan(h)), (prec, r, w, h)
            fin = ~torch.isnan(w)
            assert torch.equal(w[fin], h[fin]), (prec, r, w, h)      # +-inf with the right sign
        others = [i for i in range(n) if i not in rows]
        tol = 2e-2 if prec == _capi.PRECISION_BF16 else 1e-5
        scale = g[others].norm(dim=-1)[None, :] * t.norm(dim=-1)[:, None]
        assert ((got[:, others] - want[:, others]).abs() / scale).max().item() <= tol
    lab = torch.from_numpy(synthetic.make_labels(n, 2, 1).astype(np.uint8)).cuda()
    res = ctx.eval_ranking(g, t, lab, 2, n, _capi.PRECISION_FP32, want_per_query=True, ranked_cap=n)
    order0 = res["ranked_idx"][0].cpu().numpy()         # query 0: t[0,7] = 1.5 > 0
    assert order0[0] == 5 and int(res["ranked_len"][0]) == n - 2        # +inf first; rows 9 and 12 are NaN: dropped
    assert order0[n - 3] == 6                                            # -inf last


def test_edge_cases_and_errors(ctx):
    w = synthetic.dyadic_grid(300, 64, 3, 2, seed=1)
    g, t = w.gallery.cuda(), w.queries.cuda()
    # a label >= n_classes is rejected, not mis-counted
    bad = torch.full((300,), 5, dtype=torch.uint8, device="cuda")
    with pytest.raises(_capi.DvlError):
        ctx.eval_ranking(g, t, bad, 2, 10)
    with pytest.raises(_capi.DvlError):
        ctx.eval_ranking(g, t, torch.zeros(300, dtype=torch.uint8, device="cuda"), 1, 0)
    with pytest.raises(ValueError):
        dvl.eval_ranking(w.labels[:-1], g, t, "maxskew", 10)
    # a single gallery row, a single query, depth not a multiple of 64 or 8
    for (n, d, q) in [(1, 64, 1), (2, 20, 1), (259, 100, 2), (257, 52, 129)]:
        ww = synthetic.dyadic_grid(n, d, q, min(2, n), seed=n)
        want = spec.measure(ww.labels, ww.gallery, ww.queries, 1.0)
        got = dvl.evaluate(ww.labels, ww.gallery.cuda(), ww.queries.cuda(), 1.0)
        assert max(abs(a - b) for a, b in zip(_flat(got), _flat(want))) <= TIGHT, (n, d, q)
    # fp16 embeddings (CLIP on CUDA is fp16; the reference does .float() on them)
    w16 = synthetic.clip_like(5000, 128, 8, 2, seed=3)
    g16 = w16.gallery.half()
    want = spec.measure(w16.labels, g16.float(), w16.queries, 100)
    got = dvl.evaluate(w16.labels, g16.cuda(), w16.queries.cuda(), 100)
    assert max(abs(a - b) for a, b in zip(_flat(got), _flat(want))) <= METRIC_TOL


def test_measure_bias_drop_in_with_stub_encoders():
    """measure_bias end to end through the reference's signature with stub encoders: the dataset hands out
    embeddings as "images", the model's encoders are identities / a fixed table."""
    from debias_vision_lang_b200.datasets import EmbeddingGallery
    w = synthetic.clip_like(3000, 64, 319, 2, seed=21)
    raw_text = w.queries * 3.0          # un-normalised text features: get_prompt_embeddings normalises

    class Stub(torch.nn.Module):
        def encode_image(self, x):
            return x

        def encode_text(self, tok):
            return raw_text.cuda()[tok]

    tok = lambda prompts: torch.arange(len(prompts))
    got = dvl.measure_bias(Stub(), None, tok, attribute="gender",
                           opts={"dataset": EmbeddingGallery(w.gallery, w.labels), "num_workers": 0})
    t_norm = raw_text / raw_text.norm(dim=-1, keepdim=True)
    want = spec.measure(w.labels, w.gallery, t_norm, 1.0)
    assert set(got) == {"maxskew", "ndkl"} and set(got["ndkl"]) == {"eq_opp", "dem_par"}
    assert max(abs(a - b) for a, b in zip(_flat(got), _flat(want))) <= METRIC_TOL


def test_measure_bias_gallery_cache(tmp_path):
    """opts["use_cache"]: the first call encodes the gallery and writes the cache, the second is served
    from it without touching the dataset or the image encoder, with the identical result."""
    from debias_vision_lang_b200.datasets import EmbeddingGallery
    w = synthetic.clip_like(2000, 32, 319, 7, seed=22)

    class Stub(torch.nn.Module):
        image_calls = 0

        def encode_image(self, x):
            Stub.image_calls += 1
            return x

        def encode_text(self, tok):
            return w.queries.cuda()[tok]

    tok = lambda prompts: torch.arange(len(prompts))
    cache = str(tmp_path / "fairface_val_race")
    first = dvl.measure_bias(Stub(), None, tok, attribute="race",
                             opts={"dataset": EmbeddingGallery(w.gallery, w.labels, "race"), "num_workers": 0,
                                   "use_cache": cache})
    assert Stub.image_calls > 0 and EmbeddingGallery.is_cached(cache)
    calls = Stub.image_calls

    class NoDataset:
        def __len__(self):
            raise AssertionError("the dataset must not be touched when the cache is present")

    second = dvl.measure_bias(Stub(), None, tok, attribute="race",
                              opts={"dataset": NoDataset(), "num_workers": 0, "use_cache": cache})
    assert Stub.image_calls == calls
    assert _flat(first) == _flat(second)
    want = spec.measure(w.labels, w.gallery, w.queries, 1.0)
    assert max(abs(a - b) for a, b in zip(_flat(first), _flat(want))) <= METRIC_TOL


def test_full_size_properties_cfg2(ctx):
    """BASELINE config 2 at full size (108,501 x 512 x 319, C=7, k=N): size-independent properties.
    dem_par MaxSkew@N is exactly 0 and eq_opp MaxSkew@N = ln(C*max_c n_c/N); every
    ranked list is a permutation; scores along each ranked list are non-increasing; the run is
    bit-reproducible."""
    w = synthetic.baseline_workload("cfg2", device="cuda")
    n, q = w.gallery.shape[0], 32
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    t = w.queries[:q].contiguous()
    res = ctx.eval_ranking(w.gallery, t, lab, 7, n, _capi.PRECISION_FP32, want_per_query=True, ranked_cap=n,
                           want_scores=True)
    pq = res["per_query"].cpu().numpy()
    counts = np.bincount(w.labels, minlength=7)
    assert np.all(pq[:, 1] == 0.0)
    assert np.abs(pq[:, 0] - math.log(7 * counts.max() / n)).max() <= 1e-12
    assert torch.all(res["ranked_len"] == n)
    idx = res["ranked_idx"].long()
    assert torch.equal(torch.sort(idx, dim=1).values, torch.arange(n, device="cuda")[None, :].expand(q, n))
    ranked_scores = torch.gather(res["scores"], 1, idx)
    assert torch.all(ranked_scores[:, 1:] <= ranked_scores[:, :-1])
    # ties broken by lowest index
    tie = ranked_scores[:, 1:] == ranked_scores[:, :-1]
    assert torch.all(idx[:, 1:][tie] > idx[:, :-1][tie])
    res2 = ctx.eval_ranking(w.gallery, t, lab, 7, n, _capi.PRECISION_FP32, want_per_query=True)
    assert torch.equal(res2["per_query"], res["per_query"])
    # NDKL of 4 sampled queries against the oracle on the CUDA path's own scores
    want, _ = spec.eval_scores_matrix(res["scores"][:4].cpu().numpy(), w.labels, n)
    assert np.abs(pq[:4] - want).max() <= TIGHT


def test_clip_logits_both_sides_normalised():
    """DebiasCLIP.forward's logits (model/model.py:276-289): both sides normalised, times logit_scale."""
    torch.manual_seed(8)
    img = torch.randn(700, 512, device="cuda") * 5
    txt = torch.randn(9, 512, device="cuda") * 2
    scale = 100.0
    want = scale * (img / img.norm(dim=-1, keepdim=True)).double() @ (txt / txt.norm(dim=-1, keepdim=True)).double().T
    per_image, per_text = dvl.clip_logits(img, txt, scale)
    assert per_image.shape == (700, 9) and per_text.shape == (9, 700)
    assert (per_image.double() - want).abs().max().item() <= scale * 1e-5
    assert torch.equal(per_text, per_image.t())


def test_adversary_input_is_contiguous_logits_per_image():
    """the adversary's input (model/model.py:337-361) = logits_per_image of DebiasCLIP.forward
    (model/model.py:276-289), image-major and contiguous, against the reference formula in float64."""
    torch.manual_seed(5)
    img = torch.randn(3001, 512, device="cuda") * 3.0
    txt = torch.randn(319, 512, device="cuda")
    scale = float(np.exp(4.6052))
    x = measuring_bias.adversary_input(img, txt, scale)
    assert x.shape == (3001, 319) and x.is_contiguous() and x.dtype == torch.float32
    i64, t64 = img.double(), txt.double()
    want = scale * (i64 / i64.norm(dim=-1, keepdim=True)) @ (t64 / t64.norm(dim=-1, keepdim=True)).t()
    assert (x.double() - want).abs().max().item() <= scale * 1e-5
    per_image, _ = measuring_bias.clip_logits(img, txt, scale)
    assert (x - per_image).abs().max().item() <= scale * 2e-6          # same numbers as the query-major GEMM
    # it feeds an Adversary-shaped MLP as is
    mlp = torch.nn.Sequential(torch.nn.Linear(319, 32), torch.nn.ReLU(), torch.nn.Linear(32, 1)).cuda()
    assert torch.sigmoid(mlp(x)).shape == (3001, 1)


def test_graph_replay_matches_direct_launches(ctx):
    """The full-sort route replays a captured CUDA graph from the third identical call on.  Replays must
    see new contents of the same buffers, survive a differently shaped call in between (workspace
    growth invalidates the graph) and keep per-stage profiling working."""
    w = synthetic.clip_like(6000, 64, 9, 3, seed=77, device="cuda")
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    g, t = w.gallery.clone(), w.queries.clone()
    out_pq = []
    for _ in range(4):                                   # direct, capture, replay, replay
        r = ctx.eval_ranking(g, t, lab, 3, 6000, _capi.PRECISION_FP32, want_per_query=True, ranked_cap=6000)
        out_pq.append((r["per_query"].clone(), r["ranked_idx"].clone(), list(r["mean"])))
    for pq, idx, mean in out_pq[1:]:
        assert torch.equal(pq, out_pq[0][0]) and torch.equal(idx, out_pq[0][1]) and mean == out_pq[0][2]
    want = spec.measure(w.labels, w.gallery.cpu(), w.queries.cpu(), 1.0)
    assert np.abs(np.asarray(out_pq[0][2]) - np.asarray(_flat(want))).max() <= METRIC_TOL
    # same buffers, new contents: a replay must not return the old answer
    g.copy_(torch.flip(w.gallery, dims=[0]))
    lab2 = torch.flip(lab, dims=[0]).contiguous()
    lab.copy_(lab2)
    r2 = ctx.eval_ranking(g, t, lab, 3, 6000, _capi.PRECISION_FP32, want_per_query=True, ranked_cap=6000)
    want2 = spec.measure(w.labels[::-1].copy(), torch.flip(w.gallery, dims=[0]).cpu(), w.queries.cpu(), 1.0)
    assert np.abs(np.asarray(r2["mean"]) - np.asarray(_flat(want2))).max() <= METRIC_TOL
    # a larger call in between grows the workspace; the small call must still be right afterwards
    big = synthetic.clip_like(20000, 64, 12, 3, seed=78, device="cuda")
    ctx.eval_ranking(big.gallery, big.queries, torch.from_numpy(big.labels.astype(np.uint8)).cuda(), 3, 20000,
                     _capi.PRECISION_FP32)
    for _ in range(3):
        r3 = ctx.eval_ranking(g, t, lab, 3, 6000, _capi.PRECISION_FP32, want_per_query=True, ranked_cap=6000)
        assert torch.equal(r3["per_query"], r2["per_query"]) and torch.equal(r3["ranked_idx"], r2["ranked_idx"])
    # profiling inside replays
    ctx.get_profile()
    ctx.set_profiling(True)
    try:
        for _ in range(4):
            r4 = ctx.eval_ranking(g, t, lab, 3, 6000, _capi.PRECISION_FP32, want_per_query=True, ranked_cap=6000)
            assert torch.equal(r4["per_query"], r2["per_query"])
        prof = ctx.get_profile()
    finally:
        ctx.set_profiling(False)
    # four profiled calls, each: one GEMM, one partition launch of the bucket sort (its "scatter" stage)
    assert prof["sort_scatter"][0] > 0 and prof["sort_scatter"][1] == 4 and prof["gemm"][1] == 4


def test_cta_pair_gemm_variant_passes_the_parity_suite():
    """The similarity GEMM has a CTA-pair variant (tcgen05.mma.cta_group::2, clusters of two CTAs) that the
    library only picks for the fp32 mode on large query counts.  DVL_GEMM_PAIR=1 forces it for every shape
    (the switch is read once per process, hence the subprocess): similarity tolerances, exactness on the
    dyadic grid, the fp32 range, fused normalisation, the golden vectors end to end and the streaming
    top-k route must all hold with it."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DVL_GEMM_PAIR="1")
    sel = ("similarity or golden or fused_normalisation or stage_isolation or streaming_equals_full_sort "
           "or pilot_threshold_route or sharded_stages")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests"), "-m", "gpu", "-x", "-q",
                        "-k", f"({sel}) and not cta_pair", "-p", "no:cacheprovider"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
