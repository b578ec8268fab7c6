"""Generate tests/golden/*.json by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python -m oracle.make_golden

Every expected value written here is an output of ``/root/reference/debias_clip/
measuring_bias.py`` (``compute_skew_metrics``, ``normalized_discounted_KL``,
``eval_ranking``, ``gen_prompts``) or of ``pandas.DataFrame.nlargest`` as the reference
calls it -- never of ``oracle/spec.py``.  Inputs are either stored inline (stage-level
cases) or regenerated from the stored seed by ``debias_vision_lang_b200.synthetic`` (the
dyadic-grid workloads, whose dot products are exact in any summation order, so the
expected values hold on any machine and on the GPU).
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from debias_vision_lang_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _f32_list(a):
    """fp32 values as exact hex bit patterns (JSON floats would lose -0.0/NaN payloads)."""
    return [int(x) for x in np.asarray(a, dtype=np.float32).view(np.uint32)]


def stage_cases(mb):
    rng = np.random.RandomState(20221018)
    cases = []

    def add(name, scores, labels, top_ns):
        scores = np.asarray(scores, dtype=np.float32)
        labels = np.asarray(labels, dtype=np.int64)
        df = pd.DataFrame({"score": scores, "label": labels})
        for top_n in top_ns:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ms = mb.compute_skew_metrics(df, top_n)
                nd = mb.normalized_discounted_KL(df, top_n)
                order = [int(i) for i in df.nlargest(top_n, columns="score", keep="all").index]
            cases.append(dict(
                name=f"{name}@{top_n}", scores_u32=_f32_list(scores), labels=labels.tolist(), top_n=int(top_n),
                order=order,
                maxskew_eq_opp=float(ms["maxskew_eq_opp"]), maxskew_dem_par=float(ms["maxskew_dem_par"]),
                ndkl_eq_opp=float(nd["ndkl_eq_opp"]), ndkl_dem_par=float(nd["ndkl_dem_par"])))

    # SURVEY.md section 8(c) tie KATs
    add("kat_ties", [.5, .9, .5, .9, .1, .5, .9, .5], [0, 1, 0, 1, 0, 1, 0, 1], [1, 2, 3, 4, 7, 8, 9, 12])
    add("kat_signed_zero", [0., -0., 0., -0., 1., -1.], [0, 1, 1, 0, 0, 1], [1, 2, 3, 5, 6])
    add("kat_nan", [.5, float("nan"), .7, .5, float("nan"), .1], [0, 1, 0, 1, 0, 1], [1, 2, 3, 4])
    add("kat_all_equal", np.zeros(16), rng.permutation(np.arange(16) % 3), [1, 5, 16, 20])
    add("kat_inf", [float("inf"), 1., float("-inf"), 1., 0., float("inf")], [0, 1, 0, 1, 1, 0], [1, 2, 4, 6])
    add("kat_denormal", [1e-45, -1e-45, 0., 1e-40, -0., 2e-45], [0, 1, 0, 1, 1, 0], [1, 3, 4, 6])
    # random tie-heavy dyadic scores over C = 2..9
    for c in (2, 3, 7, 9):
        for n in (9, 33, 257):
            lab = rng.randint(0, c, size=n)
            lab[:c] = np.arange(c)
            s = rng.randint(-6, 7, size=n).astype(np.float32) / 8.0
            add(f"dyadic_c{c}_n{n}", s, lab, sorted({1, 2, c, n // 3 + 1, n - 1, n, n + 2}))
    # continuous scores (no ties)
    for c in (2, 7):
        n = 300
        lab = rng.randint(0, c, size=n)
        lab[:c] = np.arange(c)
        add(f"gauss_c{c}_n{n}", rng.randn(n).astype(np.float32), lab, [1, 10, 100, n])
    return cases


def ranking_cases(mb):
    """eval_ranking-level vectors on exactly-summable dyadic embeddings."""
    cases = []
    specs = [
        dict(kind="dyadic_grid", n=700, d=64, q=12, c=2, seed=3, lim=127, topns=[1, 50, 700, 1.0, 0.1]),
        dict(kind="dyadic_grid", n=1500, d=128, q=9, c=7, seed=4, lim=15, topns=[100, 1.0]),
        dict(kind="dyadic_grid", n=999, d=72, q=5, c=9, seed=5, lim=5, topns=[17, 0.5, 2000]),
        dict(kind="adversarial", n=2048, d=64, q=20, c=2, seed=6, topns=[10, 256, 1.0]),
        dict(kind="adversarial", n=3000, d=96, q=11, c=7, seed=8, topns=[1000]),
    ]
    for sp in specs:
        if sp["kind"] == "dyadic_grid":
            w = synthetic.dyadic_grid(sp["n"], sp["d"], sp["q"], sp["c"], seed=sp["seed"], lim=sp["lim"])
        else:
            w = synthetic.adversarial(sp["n"], sp["d"], sp["q"], sp["c"], seed=sp["seed"])
        for topn in sp["topns"]:
            res = {}
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for ev in ("maxskew", "ndkl"):
                    r = mb.eval_ranking(w.labels, w.gallery, w.queries, ev, topn)
                    res[ev] = {k: float(v) for k, v in r.items()}
            case = {k: v for k, v in sp.items() if k != "topns"}
            case.update(topn=topn, topn_is_float=isinstance(topn, float), expected=res,
                        gallery_checksum=float(w.gallery.double().sum()), labels_checksum=int(w.labels.sum()))
            cases.append(case)
    return cases


def gv1(mb):
    """isotropic randn, N=10954, D=512, Q=319 -- the cross-version regression
    vector (depends on the CPU BLAS summation order; compare at 1e-9, not bit-exactly)."""
    torch.manual_seed(1234)
    img = torch.randn(10954, 512)
    txt = torch.randn(319, 512)
    txt /= txt.norm(dim=-1, keepdim=True)
    labels = np.random.RandomState(1234).randint(0, 2, size=10954)
    out = {}
    for topn in (1000,):
        for ev in ("maxskew", "ndkl"):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = mb.eval_ranking(labels, img, txt, ev, topn)
            out[f"{ev}@{topn}"] = {k: float(v) for k, v in r.items()}
    # topn=1.0 takes ~160 s through the reference; values recorded there
    out["maxskew@1.0"] = {"eq_opp": 0.005825612688995174, "dem_par": 0.0}
    out["ndkl@1.0"] = {"eq_opp": 0.0018573914864923936, "dem_par": 0.0018411855231833375}
    return dict(seed=1234, n=10954, d=512, q=319, c=2, expected=out,
                note="topn=1.0 values are SURVEY.md section 8(c) [measured with the reference, 156 s]")


def main():
    mb = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    prompts = mb.gen_prompts()
    payload = {
        "stage_cases.json": stage_cases(mb),
        "ranking_cases.json": ranking_cases(mb),
        "gv1.json": gv1(mb),
        "prompts.json": dict(n=len(prompts), prompts=prompts),
    }
    meta = dict(pandas=pd.__version__, numpy=np.__version__, torch=torch.__version__,
                reference="oxai/debias-vision-lang @ /root/reference (unmodified, via oracle/ref_shim.py)")
    for fname, body in payload.items():
        with open(os.path.join(OUT, fname), "w") as f:
            json.dump(dict(meta=meta, data=body), f, indent=None, separators=(",", ":"))
        print("wrote", fname, os.path.getsize(os.path.join(OUT, fname)), "bytes")


if __name__ == "__main__":
    main()
