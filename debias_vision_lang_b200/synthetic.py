"""Seeded synthetic galleries, query sets and attribute labels in the shapes of
BASELINE.json's configs.  There is no network, so there are no
FairFace images or CLIP weights; these stand in for what the encoders
(``measuring_bias.py:75-99`` of the reference) would hand to the hot path.

Two families:

* ``clip_like``  -- raw (un-normalised, F1) image embeddings with varying norms and a weak
  label<->similarity signal, unit-norm prompt-array queries (11 templates x 29 concepts,
  tiled up to Q);
* ``dyadic_grid`` / ``adversarial`` -- every component is k/64 with |k| <= 127, so each
  product and every partial sum of a D <= 1024 dot product is exactly representable in
  fp32: similarities are identical whatever the summation order (CPU BLAS, tensor
  cores), which makes rankings comparable bit for bit and produces many exact ties.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

N_TEMPLATES = 11
N_CONCEPTS = 29

CLASS_PRIORS = {
    2: [0.53, 0.47],                                               # gender
    7: [0.19, 0.13, 0.11, 0.14, 0.14, 0.15, 0.14],                 # race
    9: [0.02, 0.12, 0.11, 0.30, 0.22, 0.12, 0.07, 0.03, 0.01],     # age
}


@dataclass
class Workload:
    name: str
    gallery: torch.Tensor      # [N, D] fp32, raw image embeddings
    queries: torch.Tensor      # [Q, D] fp32, unit rows
    labels: np.ndarray         # [N] int64 in 0..C-1, every class present
    n_classes: int
    topn: object               # int, or float proportion


def make_labels(n: int, n_classes: int, seed: int) -> np.ndarray:
    rng = np.random.RandomState(seed)
    pri = CLASS_PRIORS.get(n_classes)
    if pri is None:
        pri = np.full(n_classes, 1.0 / n_classes)
    pri = np.asarray(pri, dtype=np.float64)
    lab = rng.choice(n_classes, size=n, p=pri / pri.sum()).astype(np.int64)
    # the reference asserts every class is present (datasets.py:43)
    lab[:n_classes] = np.arange(n_classes)
    return lab


def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True)


def clip_like(n: int, d: int, q: int, n_classes: int, seed: int = 1234,
              device: str = "cpu", chunk: int = 1 << 18) -> Workload:
    """g_i = r_i * unit(a*mu_img + beta*v_{label_i} + z_i); t_q = unit(a*mu_txt + u_template
    + 0.3*w_concept + z_q); cos(mu_img, mu_txt) ~ 0.25."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def rn(*shape):
        return torch.randn(*shape, generator=gen, device=device, dtype=torch.float32)

    labels = make_labels(n, n_classes, seed)
    lab_t = torch.from_numpy(labels).to(device)
    mu_img = _unit(rn(d))
    mu_txt = _unit(0.25 * mu_img + (1 - 0.25 ** 2) ** 0.5 * _unit(rn(d) - (rn(d) @ mu_img) * mu_img))
    v = _unit(rn(n_classes, d))
    gallery = torch.empty(n, d, dtype=torch.float32, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        z = rn(hi - lo, d) / d ** 0.5
        body = _unit(1.2 * mu_img[None, :] + 0.15 * v[lab_t[lo:hi]] + z)
        r = 10.0 * torch.exp(0.05 * rn(hi - lo, 1))
        gallery[lo:hi] = r * body
    u = _unit(rn(N_TEMPLATES, d))
    w = _unit(rn(N_CONCEPTS, d))
    qi = torch.arange(q, device=device)
    tmpl = (qi // N_CONCEPTS) % N_TEMPLATES
    conc = qi % N_CONCEPTS
    zq = rn(q, d) / d ** 0.5 * 0.3
    # a faint coupling of some concepts to the class directions keeps the metrics non-trivial
    bias = 0.08 * v[conc % n_classes]
    queries = _unit(1.2 * mu_txt[None, :] + u[tmpl] + 0.3 * w[conc] + bias + zq)
    return Workload(f"clip_like_{n}x{d}x{q}_c{n_classes}", gallery, queries, labels, n_classes, 1.0)


GEN_CHUNK = 65536          # rows per independently seeded block of a `clip_like_rows` gallery


def clip_like_rows(n_total: int, d: int, q: int, n_classes: int, seed: int, lo: int = 0, hi: Optional[int] = None,
                   device: str = "cpu", dtype: torch.dtype = torch.float32) -> Workload:
    """Rows [lo, hi) of a `clip_like` gallery of `n_total` rows whose content does not depend on how it is
    cut: every block of GEN_CHUNK rows (and its labels) has its own seed, the shared directions and the
    queries come from the header seed.  Rank r of a gallery-sharded run generates only its own row block
    and gets exactly the rows a single GPU holding the whole gallery would have there -- which is what
    lets `bench.py` compare the per-query results of an N-GPU run with the one-GPU ones bit for bit.
    (Same generator family as `clip_like`; bit-identical across runs on the same device type only.)"""
    hi = n_total if hi is None else hi
    assert 0 <= lo <= hi <= n_total
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def rn(*shape, g=gen):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32)

    mu_img = _unit(rn(d))
    mu_txt = _unit(0.25 * mu_img + (1 - 0.25 ** 2) ** 0.5 * _unit(rn(d) - (rn(d) @ mu_img) * mu_img))
    v = _unit(rn(n_classes, d))
    u = _unit(rn(N_TEMPLATES, d))
    w = _unit(rn(N_CONCEPTS, d))
    qi = torch.arange(q, device=device)
    tmpl = (qi // N_CONCEPTS) % N_TEMPLATES
    conc = qi % N_CONCEPTS
    zq = rn(q, d) / d ** 0.5 * 0.3
    bias = 0.08 * v[conc % n_classes]
    queries = _unit(1.2 * mu_txt[None, :] + u[tmpl] + 0.3 * w[conc] + bias + zq).to(dtype)

    pri = np.asarray(CLASS_PRIORS.get(n_classes, np.full(n_classes, 1.0 / n_classes)), dtype=np.float64)
    gallery = torch.empty(hi - lo, d, dtype=dtype, device=device)
    labels = np.empty(hi - lo, dtype=np.int64)
    cgen = torch.Generator(device=device)
    for c in range(lo // GEN_CHUNK, -(-hi // GEN_CHUNK) if hi > lo else 0):
        c_lo, c_hi = c * GEN_CHUNK, min(n_total, (c + 1) * GEN_CHUNK)
        rows = c_hi - c_lo
        lab_c = np.random.RandomState((seed * 1000003 + c) % (2 ** 31)).choice(n_classes, size=rows, p=pri / pri.sum())
        if c == 0:
            lab_c[:n_classes] = np.arange(n_classes)          # every class present (datasets.py:43)
        cgen.manual_seed(seed * 7919 + 104729 * (c + 1))
        z = rn(rows, d, g=cgen) / d ** 0.5
        body = _unit(1.2 * mu_img[None, :] + 0.15 * v[torch.from_numpy(lab_c).to(device)] + z)
        r = 10.0 * torch.exp(0.05 * rn(rows, 1, g=cgen))
        a, b = max(lo, c_lo), min(hi, c_hi)
        gallery[a - lo:b - lo] = (r * body)[a - c_lo:b - c_lo].to(dtype)
        labels[a - lo:b - lo] = lab_c[a - c_lo:b - c_lo]
    return Workload(f"clip_like_rows_{n_total}x{d}x{q}_c{n_classes}[{lo}:{hi}]", gallery, queries, labels, n_classes, 1.0)


def dyadic_grid(n: int, d: int, q: int, n_classes: int, seed: int = 7, lim: int = 127,
                device: str = "cpu") -> Workload:
    """All components k/64, |k| <= lim <= 127: order-independent exact fp32 dot products
    for D <= 1024 (|sum| <= 1024*127*127 < 2^24 in units of 2^-12)."""
    assert d <= 1024 and lim <= 127
    rng = np.random.RandomState(seed)
    g = rng.randint(-lim, lim + 1, size=(n, d)).astype(np.float32) / 64.0
    t = rng.randint(-lim, lim + 1, size=(q, d)).astype(np.float32) / 64.0
    labels = make_labels(n, n_classes, seed + 1)
    return Workload(f"dyadic_{n}x{d}x{q}_c{n_classes}", torch.from_numpy(g).to(device),
                    torch.from_numpy(t).to(device), labels, n_classes, 1.0)


def adversarial(n: int, d: int, q: int, n_classes: int, seed: int = 11, dup_frac: float = 0.02,
                n_shards: int = 8, device: str = "cpu") -> Workload:
    """BASELINE config 5's adversarial prompt-array batch on the dyadic grid: >= dup_frac of the
    gallery rows are exact duplicates of rows living on *other* shards (cross-shard exact
    ties), queries come in near-duplicate families (same template vector, a few components
    changed), query 0 is all zeros (every score ties: m = N overflow, with +0/-0 mixed),
    and a small-range grid makes ties frequent everywhere."""
    rng = np.random.RandomState(seed)
    lim = 3                                   # tiny alphabet => massive ties
    g = rng.randint(-lim, lim + 1, size=(n, d)).astype(np.float32) / 64.0
    n_dup = max(1, int(n * dup_frac))
    shard = -(-n // n_shards)
    src = rng.randint(0, n, size=n_dup)
    dst = (src + shard * rng.randint(1, max(2, n_shards), size=n_dup) + rng.randint(0, shard, size=n_dup)) % n
    g[dst] = g[src]
    base = rng.randint(-lim, lim + 1, size=(max(1, q // 8 + 1), d)).astype(np.float32) / 64.0
    t = base[np.arange(q) // 8].copy()
    for i in range(q):                        # near-duplicates: perturb (i % 8) components
        for j in range(i % 8):
            t[i, (17 * i + 31 * j) % d] = rng.randint(-lim, lim + 1) / 64.0
    t[0] = 0.0
    if q > 1:
        t[1] = -t[2] if q > 2 else t[1]        # mirrored pair: descending order reversed
    labels = make_labels(n, n_classes, seed + 1)
    # correlate index order with the label (F10): first half mostly class 0
    labels[: n // 2][rng.rand(n // 2) < 0.6] = 0
    labels[:n_classes] = np.arange(n_classes)
    return Workload(f"adversarial_{n}x{d}x{q}_c{n_classes}", torch.from_numpy(g).to(device),
                    torch.from_numpy(t).to(device), labels, n_classes, 1.0)


# name -> (N, D, Q, C, topn, precision) of BASELINE.json's five configs
BASELINE_CONFIGS = {
    "cfg1": dict(n=10_954, d=512, q=319, c=2, topn=1000, precision="fp32"),
    "cfg2": dict(n=108_501, d=512, q=319, c=7, topn=1.0, precision="fp32"),
    "cfg3": dict(n=1_000_000, d=1024, q=1000, c=2, topn=1000, precision="bf16"),
    "cfg4": dict(n=10_000_000, d=768, q=10_000, c=2, topn=1000, precision="bf16"),
    "cfg5": dict(n=1_000_000, d=512, q=1000, c=2, topn=1000, precision="fp32"),
}


def baseline_workload(name: str, device: str = "cpu", n_override: Optional[int] = None,
                      q_override: Optional[int] = None) -> Workload:
    cfg = BASELINE_CONFIGS[name]
    n = n_override or cfg["n"]
    q = q_override or cfg["q"]
    seed = 1234 + int(name[-1])
    if name == "cfg5":
        w = adversarial(n, cfg["d"], q, cfg["c"], seed=seed, device=device)
    else:
        w = clip_like(n, cfg["d"], q, cfg["c"], seed=seed, device=device)
    w.name = name
    w.topn = cfg["topn"]
    return w
