"""Diagnostic driver for the GPU box (not a pytest): prints error patterns stage by stage so a
wrong descriptor / barrier shows up as a recognisable pattern in one gpurun call.

    timeout 300 python tests/gpu_debug.py [stage ...]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from debias_vision_lang_b200 import _capi, synthetic  # noqa: E402
from oracle import spec  # noqa: E402


def stage_similarity(ctx):
    torch.manual_seed(0)
    for (n, d, q) in [(256, 64, 128), (300, 64, 100), (1000, 128, 130), (5000, 512, 319), (4096, 768, 256),
                      (777, 200, 33)]:
        g = torch.randn(n, d, device="cuda")
        t = torch.randn(q, d, device="cuda")
        t = t / t.norm(dim=-1, keepdim=True)
        ref64 = (t.double() @ g.double().T)
        for prec, name, tol in ((_capi.PRECISION_BF16, "bf16", 2e-2), (_capi.PRECISION_FP32, "fp32", 1e-5)):
            out = ctx.similarity(g, t, prec)
            torch.cuda.synchronize()
            if prec == _capi.PRECISION_BF16:
                want = (t.bfloat16().double() @ g.bfloat16().double().T)
            else:
                want = ref64
            scale = (g.double().norm(dim=-1)[None, :] * t.double().norm(dim=-1)[:, None])
            err = ((out.double() - want).abs() / scale)
            bad = (err > tol).nonzero()
            print(f"similarity n={n} d={d} q={q} {name}: max normwise err {err.max().item():.3e} "
                  f"(vs exact {(((out.double() - ref64).abs() / scale).max().item()):.3e}) bad={bad.shape[0]}")
            if bad.shape[0]:
                rows = torch.unique(bad[:, 0])[:16].tolist()
                cols = torch.unique(bad[:, 1])[:32].tolist()
                print("   first bad rows", rows, "cols", cols)
                print("   got ", out[bad[0, 0], bad[0, 1]].item(), "want", want[bad[0, 0], bad[0, 1]].item())


def stage_rank(ctx):
    rng = np.random.RandomState(1)
    for (n, q, c, top_n, ties) in [(100, 3, 2, 10, True), (5000, 7, 7, 5000, True), (5000, 7, 7, 100, False),
                                   (10954, 5, 2, 1000, False), (4097, 4, 9, 4097, True), (300, 2, 3, 400, True)]:
        if ties:
            s = (rng.randint(-40, 41, size=(q, n)) / 16.0).astype(np.float32)
        else:
            s = rng.randn(q, n).astype(np.float32)
        lab = synthetic.make_labels(n, c, 5)
        want, orders = spec.eval_scores_matrix(s, lab, top_n)
        cap = max(len(o) for o in orders)
        res = ctx.rank_and_score(torch.from_numpy(s).cuda(), torch.from_numpy(lab.astype(np.uint8)).cuda(), c, top_n,
                                 ranked_cap=cap)
        got = res["per_query"].cpu().numpy()
        lens = res["ranked_len"].cpu().numpy()
        idx = res["ranked_idx"].cpu().numpy()
        ok_len = all(lens[i] == len(orders[i]) for i in range(q))
        ok_idx = all(np.array_equal(idx[i, :lens[i]], orders[i]) for i in range(q)) if ok_len else False
        print(f"rank n={n} q={q} c={c} top_n={top_n} ties={ties}: max metric err {np.abs(got - want).max():.3e} "
              f"len_ok={ok_len} order_ok={ok_idx} mean_err {np.abs(np.array(res['mean']) - want.mean(0)).max():.3e}")
        if not ok_len:
            print("   lens", lens[:8], [len(o) for o in orders][:8])


def stage_e2e(ctx):
    for w, topn, prec in [(synthetic.dyadic_grid(3000, 64, 20, 2, seed=3), 100, _capi.PRECISION_FP32),
                          (synthetic.dyadic_grid(3000, 64, 20, 7, seed=4, lim=15), 3000, _capi.PRECISION_FP32),
                          (synthetic.adversarial(4096, 64, 24, 2, seed=6), 256, _capi.PRECISION_FP32),
                          (synthetic.clip_like(20000, 512, 64, 7, seed=9), 20000, _capi.PRECISION_FP32),
                          (synthetic.clip_like(20000, 512, 64, 2, seed=9), 1000, _capi.PRECISION_BF16)]:
        want = spec.measure(w.labels, w.gallery, w.queries, topn)
        res = ctx.eval_ranking(w.gallery.cuda(), w.queries.cuda(), torch.from_numpy(w.labels.astype(np.uint8)).cuda(),
                               w.n_classes, topn, prec)
        wv = [want["maxskew"]["eq_opp"], want["maxskew"]["dem_par"], want["ndkl"]["eq_opp"], want["ndkl"]["dem_par"]]
        print(f"e2e {w.name} topn={topn} prec={prec}: got {res['mean']} want {wv} "
              f"maxerr {max(abs(a - b) for a, b in zip(res['mean'], wv)):.3e}")


def stage_time(ctx):
    w = synthetic.baseline_workload("cfg2", device="cuda")
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    for prec in (_capi.PRECISION_FP32, _capi.PRECISION_BF16):
        for _ in range(2):
            ctx.eval_ranking(w.gallery, w.queries, lab, w.n_classes, w.gallery.shape[0], prec)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            r = ctx.eval_ranking(w.gallery, w.queries, lab, w.n_classes, w.gallery.shape[0], prec)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 3
        print(f"cfg2 prec={prec}: {dt * 1e3:.3f} ms/step, {w.gallery.shape[0] * w.queries.shape[0] / dt:.3e} pairs/s", r["mean"])


if __name__ == "__main__":
    stages = sys.argv[1:] or ["similarity", "rank", "e2e", "time"]
    print(torch.cuda.get_device_name(0), "lib", _capi.LIB_PATH)
    ctx = _capi.Context(0)
    for s in stages:
        t0 = time.time()
        try:
            globals()["stage_" + s](ctx)
        except Exception as e:  # keep going: later stages may still be informative
            print(f"stage {s} FAILED: {type(e).__name__}: {e}")
        print(f"--- stage {s} done in {time.time() - t0:.1f}s, launches so far {ctx.launch_count}", flush=True)
