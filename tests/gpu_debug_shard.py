"""Diagnostic (not a pytest): the NCCL test's adversarial case through the emulated sharded routes on one GPU."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from debias_vision_lang_b200 import _capi, sharded, synthetic
from oracle import spec

w = synthetic.adversarial(40_000, 64, 24, 2, seed=21, n_shards=2)
topn = 300
scores = np.stack([spec.similarity_f32(w.gallery, w.queries[i]) for i in range(24)])
want_pq, ranked = spec.eval_scores_matrix(scores, w.labels, topn)
g, t = w.gallery.cuda(), w.queries.cuda()
lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
ctxs = [_capi.Context(0) for _ in range(2)]
single = ctxs[0].eval_ranking(g, t, lab, 2, topn, _capi.PRECISION_FP32, _capi.FLAG_FULL_SORT, want_per_query=True, ranked_cap=400)
print("single vs oracle", np.abs(single["per_query"].cpu().numpy() - want_pq).max(axis=1))
stream = ctxs[0].eval_ranking(g, t, lab, 2, topn, _capi.PRECISION_FP32, _capi.FLAG_STREAM_TOPK, want_per_query=True, ranked_cap=400)
print("stream vs oracle", np.abs(stream["per_query"].cpu().numpy() - want_pq).max(axis=1))
for lean in (False, True):
    for length in (None, 2048):
        if not lean and length:
            continue
        res = sharded.evaluate_sharded_emulated(ctxs, g, t, lab, 2, topn, ranked_cap=400, lean=lean, length=length)
        d = np.abs(res["per_query"].cpu().numpy() - want_pq).max(axis=1)
        print("lean" if lean else "stages", length, "n_flagged", res["n_flagged"], "list_len", res["list_len"], "bad queries", np.nonzero(d > 1e-9)[0], d.max())
        lens = res["ranked_len"].cpu().numpy()
        for qi in np.nonzero(d > 1e-9)[0]:
            print("   q", qi, "len got", lens[qi], "want", len(ranked[qi]), "single", int(single["ranked_len"][qi]))
            gi = res["ranked_idx"][qi].cpu().numpy(); wi = ranked[qi]
            k = min(len(wi), 400, lens[qi]); diff = np.nonzero(gi[:k] != wi[:k])[0]
            print("   first diff pos", diff[:5], "got", gi[diff[:5]] if len(diff) else None, "want", wi[diff[:5]] if len(diff) else None)
