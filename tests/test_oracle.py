"""The oracle (oracle/spec.py) against every golden vector produced by the reference, and
against the live reference when it is mounted.  CPU only."""
import math
import warnings

import numpy as np
import pytest

from conftest import build_ranking_workload, load_golden, u32_to_f32
from oracle import ref_shim, spec

TOL = 1e-12


def test_stage_cases_match_reference(stage_cases):
    assert len(stage_cases) > 100
    for case in stage_cases:
        s = u32_to_f32(case["scores_u32"])
        lab = np.asarray(case["labels"])
        top_n = case["top_n"]
        assert spec.rank_indices(s, top_n).tolist() == case["order"], case["name"]
        ms = spec.eval_query(s, lab, "maxskew", top_n)
        nd = spec.eval_query(s, lab, "ndkl", top_n)
        for got, key in ((ms["eq_opp"], "maxskew_eq_opp"), (ms["dem_par"], "maxskew_dem_par"),
                         (nd["eq_opp"], "ndkl_eq_opp"), (nd["dem_par"], "ndkl_dem_par")):
            want = case[key]
            if math.isnan(want):
                assert math.isnan(got), (case["name"], key)
            else:
                assert abs(got - want) <= TOL * max(1.0, abs(want)), (case["name"], key, got, want)


def test_ranking_cases_match_reference(ranking_cases):
    for case in ranking_cases:
        w, topn = build_ranking_workload(case)
        got = spec.measure(w.labels, w.gallery, w.queries, topn)
        for ev in ("maxskew", "ndkl"):
            for k in ("eq_opp", "dem_par"):
                assert abs(got[ev][k] - case["expected"][ev][k]) <= 1e-12, (case, ev, k)


def test_gv1_regression():
    import torch
    gv = load_golden("gv1.json")
    torch.manual_seed(gv["seed"])
    img = torch.randn(gv["n"], gv["d"])
    txt = torch.randn(gv["q"], gv["d"])
    txt /= txt.norm(dim=-1, keepdim=True)
    labels = np.random.RandomState(gv["seed"]).randint(0, gv["c"], size=gv["n"])
    for key, want in gv["expected"].items():
        ev, topn = key.split("@")
        topn = float(topn) if "." in topn else int(topn)
        got = spec.eval_ranking(labels, img, txt, ev, topn)
        for k in want:
            # depends on the CPU BLAS summation order through near-tied ranks: 1e-7, not bit-exact
            assert abs(got[k] - want[k]) <= 1e-7, (key, k, got[k], want[k])


def test_topn_resolution():
    assert spec.resolve_topn(1.0, 10954) == 10954
    assert spec.resolve_topn(0.1, 10954) == 1096
    assert spec.resolve_topn(1000, 5) == 1000


def test_split_bf16x3_is_exact():
    rng = np.random.RandomState(0)
    x = (rng.randn(4096) * np.exp(rng.randn(4096) * 3)).astype(np.float32)
    b1, b2, b3 = spec.split_bf16x3(x)
    assert np.array_equal((b1.astype(np.float64) + b2 + b3).astype(np.float32), x)
    for b in (b1, b2, b3):
        assert np.all((b.view(np.uint32) & 0xFFFF) == 0)


def test_split_f16x2_error_bound():
    """Model of the default fp32 mode's operand split (two fp16 planes of the power-of-two scaled row,
    three product terms): within 3 * 2^-22 of |g||t| of the exact product over the whole fp32 range, and
    exact on the dyadic grid."""
    rng = np.random.RandomState(1)
    g = (rng.randn(300, 96) * np.exp(rng.randn(300, 1) * 20) * np.exp(rng.randn(300, 96) * 2)).astype(np.float32)
    t = (rng.randn(40, 96) * np.exp(rng.randn(40, 1) * 10)).astype(np.float32)
    g[5] = 0.0
    h0, h1, down = spec.split_f16x2(g)
    assert np.all(np.isfinite(h0.astype(np.float32))) and np.all(np.isfinite(h1.astype(np.float32)))
    rebuilt = (h0.astype(np.float64) + h1.astype(np.float64)) * down.astype(np.float64)[:, None]
    assert np.abs(rebuilt - g).max(axis=1).max() >= 0          # finite
    assert np.all(np.abs(rebuilt - g) <= 2.0 ** -22 * np.abs(g).max(axis=1, keepdims=True) + 0)
    exact = t.astype(np.float64) @ g.astype(np.float64).T
    scale = np.linalg.norm(t.astype(np.float64), axis=1)[:, None] * np.linalg.norm(g.astype(np.float64), axis=1)[None, :]
    err = np.abs(spec.similarity_split_f16x2(g, t) - exact) / np.maximum(scale, 1e-300)
    assert err.max() <= 3 * 2.0 ** -22
    k = rng.randint(-127, 128, size=(64, 128)).astype(np.float32) / 64
    q = rng.randint(-127, 128, size=(8, 128)).astype(np.float32) / 64
    assert np.array_equal(spec.similarity_split_f16x2(k, q), q.astype(np.float64) @ k.astype(np.float64).T)


@pytest.mark.needs_reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")
def test_spec_against_live_reference_random():
    import pandas as pd
    mb = ref_shim.load()
    rng = np.random.RandomState(99)
    worst = 0.0
    for trial in range(120):
        n = int(rng.randint(5, 300))
        c = int(rng.randint(2, 10))
        lab = rng.randint(0, c, size=n)
        lab[:min(c, n)] = np.arange(min(c, n))
        c_eff = len(np.unique(lab))
        lab = np.searchsorted(np.unique(lab), lab)
        s = (rng.randint(-8, 9, size=n) / 8.0).astype(np.float32) if trial % 2 else rng.randn(n).astype(np.float32)
        top_n = int(rng.randint(1, n + 3))
        df = pd.DataFrame({"score": s, "label": lab})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ms = mb.compute_skew_metrics(df, top_n)
            nd = mb.normalized_discounted_KL(df, top_n)
        a = spec.eval_query(s, lab, "maxskew", top_n)
        b = spec.eval_query(s, lab, "ndkl", top_n)
        worst = max(worst, abs(a["eq_opp"] - ms["maxskew_eq_opp"]), abs(a["dem_par"] - ms["maxskew_dem_par"]),
                    abs(b["eq_opp"] - nd["ndkl_eq_opp"]), abs(b["dem_par"] - nd["ndkl_dem_par"]))
        assert c_eff >= 1
    assert worst <= 1e-12, worst


def test_row_by_row_port_matches_golden(stage_cases, ranking_cases):
    import pandas as pd
    from oracle import port
    for case in stage_cases[::3]:
        lab = np.asarray(case["labels"])
        df = pd.DataFrame({"score": u32_to_f32(case["scores_u32"]), "label": lab})
        ms = port.maxskew_one_query(df, case["top_n"])
        assert abs(ms["eq_opp"] - case["maxskew_eq_opp"]) <= TOL and abs(ms["dem_par"] - case["maxskew_dem_par"]) <= TOL
        nd = port.ndkl_one_query(df, case["top_n"])
        if not math.isnan(case["ndkl_eq_opp"]):
            assert abs(nd["eq_opp"] - case["ndkl_eq_opp"]) <= TOL and abs(nd["dem_par"] - case["ndkl_dem_par"]) <= TOL
    case = ranking_cases[0]
    w, topn = build_ranking_workload(case)
    got = port.measure(w.labels, w.gallery, w.queries, topn)
    for ev in ("maxskew", "ndkl"):
        for k in ("eq_opp", "dem_par"):
            assert abs(got[ev][k] - case["expected"][ev][k]) <= 1e-12
