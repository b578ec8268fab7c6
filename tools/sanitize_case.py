"""Small end-to-end cases for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
Covers the bucket sort (plan / partition / sort, incl. the tie-heavy fallback), the streaming top-k with its
FILTER epilogue and select kernels, the lean sharded exchange (two emulated shards), the host-input entry
point (copy stream + events) and the device label encoding.  No result checks beyond finiteness.
(Round 2: compute-sanitizer is closed on this GPU pool -- the call answers "closed" and exits 86 -- so the case was
only run plain; correctness rests on the oracle comparisons of tests/.)"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from debias_vision_lang_b200 import _capi, measuring_bias, sharded, synthetic  # noqa: E402


def main():
    ctx = _capi.Context(0)
    w = synthetic.clip_like(9_000, 64, 6, 3, seed=3, device="cuda")
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    n = w.gallery.shape[0]
    out = []
    out.append(ctx.eval_ranking(w.gallery, w.queries, lab, 3, n, _capi.PRECISION_FP32, ranked_cap=64)["mean"])          # bucket sort
    out.append(ctx.eval_ranking(w.gallery, w.queries, lab, 3, n, _capi.PRECISION_FP32, _capi.FLAG_RADIX_SORT)["mean"])  # radix sort
    out.append(ctx.eval_ranking(w.gallery, w.queries, lab, 3, 300, _capi.PRECISION_BF16, _capi.FLAG_STREAM_TOPK)["mean"])
    ties = synthetic.adversarial(6_000, 64, 5, 2, seed=4)
    out.append(ctx.eval_ranking(ties.gallery.cuda(), ties.queries.cuda(), torch.from_numpy(ties.labels.astype(np.uint8)).cuda(),
                                2, 6_000, _capi.PRECISION_FP32)["mean"])                                                # overflow -> fallback
    ctxs = [ctx, _capi.Context(0)]
    r = sharded.evaluate_sharded_emulated(ctxs, w.gallery, w.queries, lab, 3, 200, ranked_cap=256, length=8)          # re-exchange
    out.append([r["maxskew"]["eq_opp"], r["ndkl"]["eq_opp"]])
    g, t, l = w.gallery.cpu().numpy(), w.queries.cpu().numpy(), w.labels.astype(np.uint8)
    out.append(ctx.eval_ranking_host(g, t, l, 3, n, _capi.PRECISION_FP32))
    out.append(list(measuring_bias.evaluate(w.labels, w.gallery, w.queries, 1.0)["ndkl"].values()))                     # device label encode
    torch.cuda.synchronize()
    flat = [x for row in out for x in row]
    assert all(np.isfinite(flat)), flat
    print("sanitize_case ok", len(flat), "numbers")


if __name__ == "__main__":
    main()
