"""Full-size parity on one GPU: BASELINE configs 3 and 5 at their real shapes (1 M-row galleries, 1000 queries,
top-1000) and one GPU's shard of config 4 (1.25 M x 768 x 10 k, bf16), on dyadic-grid data whose dot products
are exact in fp32 whatever the summation order (components k/64, |k| <= 127, are exact in bf16 too).  The
CUDA path ranks ALL queries; the ranked indices, ranked lengths and the four metrics of a query sample are
compared bit for bit / to 1e-9 with the oracle (oracle/spec.py on exact fp32 BLAS scores).  Also the
host-input entry points at a size where the chunked upload really has several blocks."""
import numpy as np
import pytest
import torch

from debias_vision_lang_b200 import _capi, sharded, synthetic
from oracle import spec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    c = _capi.Context(0)
    yield c
    c.close()


def _dyadic(n, d, q, seed, lim):
    rng = np.random.RandomState(seed)
    g = np.empty((n, d), dtype=np.float32)
    for lo in range(0, n, 1 << 17):                      # int8 -> float32 in blocks: bounded temporaries
        hi = min(n, lo + (1 << 17))
        g[lo:hi] = rng.randint(-lim, lim + 1, size=(hi - lo, d), dtype=np.int8)
    g /= 64.0
    t = (rng.randint(-lim, lim + 1, size=(q, d)) / 64.0).astype(np.float32)
    return g, t


def _check_sample(res, g, t, labels, top_n, sample, cap):
    scores = (t[sample] @ g.T).astype(np.float32)        # exact: every partial sum is representable
    per_query, ranked = spec.eval_scores_matrix(scores, labels, top_n)
    got_pq = res["per_query"].cpu().numpy()[sample]
    got_len = res["ranked_len"].cpu().numpy()[sample]
    got_idx = res["ranked_idx"].cpu().numpy()[sample]
    for j, qi in enumerate(sample):
        want = ranked[j]
        assert got_len[j] == len(want), (qi, got_len[j], len(want))
        shown = min(len(want), cap)
        assert np.array_equal(got_idx[j, :shown], want[:shown]), qi
    assert np.abs(got_pq - per_query).max() <= 1e-9


def test_config3_full_size_bf16_sampled_queries_exact(ctx):
    n, d, q, top_n, cap = 1_000_000, 1024, 1000, 1000, 1536
    g, t = _dyadic(n, d, q, seed=303, lim=127)            # |sum| <= 1024 * 127^2 < 2^24 in units of 2^-12
    labels = synthetic.make_labels(n, 2, 304)
    lab = torch.from_numpy(labels.astype(np.uint8)).cuda()
    res = ctx.eval_ranking(torch.from_numpy(g).cuda().bfloat16(), torch.from_numpy(t).cuda().bfloat16(), lab, 2, top_n,
                           _capi.PRECISION_BF16, 0, want_per_query=True, ranked_cap=cap)
    sample = np.random.RandomState(1).choice(q, size=16, replace=False)
    _check_sample(res, g, t, labels, top_n, sample, cap)


def test_config5_full_size_fp32_adversarial_sampled_queries_exact(ctx):
    w = synthetic.baseline_workload("cfg5")                # 1 M x 512 x 1000, the adversarial dyadic batch
    g, t = w.gallery.numpy(), w.queries.numpy()
    lab = torch.from_numpy(w.labels.astype(np.uint8)).cuda()
    top_n, cap = 1000, 4096
    res = ctx.eval_ranking(w.gallery.cuda(), w.queries.cuda(), lab, w.n_classes, top_n, _capi.PRECISION_FP32, 0,
                           want_per_query=True, ranked_cap=cap)
    # query 0 is all zeros (every row ties: m = N), query 1 mirrors query 2; plus a random sample
    sample = np.concatenate([[0, 1, 2], np.random.RandomState(2).choice(np.arange(3, 1000), size=13, replace=False)])
    _check_sample(res, g, t, w.labels, top_n, sample, cap)
    assert int(res["ranked_len"][0]) == g.shape[0]


def test_config4_shard_size_bf16_sharded_stages_sampled_queries_exact(ctx):
    """One GPU's shard of config 4 at 8 GPUs (1.25 M x 768, 10 k queries, top-1000), through the sharded stages
    with a non-zero index base (world = 1 communicator: the collectives degenerate, the kernels do not)."""
    n, d, q, top_n, cap = 1_250_048, 768, 10_000, 1000, 1280
    g, t = _dyadic(n, d, q, seed=404, lim=127)
    labels = synthetic.make_labels(n, 2, 405)
    lab = torch.from_numpy(labels.astype(np.uint8)).cuda()
    res = sharded.evaluate_sharded(ctx, sharded.SoloComm(), torch.from_numpy(g).cuda().bfloat16(),
                                   torch.from_numpy(t).cuda().bfloat16(), lab, 0, n, 2, top_n, {"precision": "bf16"},
                                   ranked_cap=cap)
    sample = np.random.RandomState(3).choice(q, size=16, replace=False)
    _check_sample(res, g, t, labels, top_n, sample, cap)
    # the same shard handed over from pinned host memory (chunked upload under the streaming top-k)
    host = sharded.evaluate_sharded_host(ctx, sharded.SoloComm(), torch.from_numpy(g).bfloat16().pin_memory(),
                                         torch.from_numpy(t).bfloat16().pin_memory(),
                                         torch.from_numpy(labels.astype(np.uint8)).pin_memory(), 0, n, 2, top_n,
                                         {"precision": "bf16"}, ranked_cap=cap)
    assert torch.equal(host["per_query"], res["per_query"])
    assert torch.equal(host["ranked_len"], res["ranked_len"])
    assert torch.equal(host["ranked_idx"], res["ranked_idx"])


@pytest.mark.parametrize("n,d,q,c,topn,prec", [
    (108_501, 512, 40, 7, 108_501, _capi.PRECISION_FP32),     # config 2 shape: full-sort route, 4 upload slices
    (300_000, 128, 33, 2, 500, _capi.PRECISION_FP32),         # streaming route fed by the upload (progressive)
    (70_001, 96, 17, 3, 1000, _capi.PRECISION_BF16),          # fp32 host rows -> bf16 planes, ragged last block
])
def test_host_entry_point_equals_device_entry_point_and_oracle(ctx, n, d, q, c, topn, prec):
    """dvl_eval_ranking_host (queries/labels first, gallery in row blocks on the copy stream, operand split and
    similarity per landed block) == dvl_eval_ranking on resident inputs, bit for bit, and == the oracle."""
    w = synthetic.dyadic_grid(n, d, q, c, seed=n % 97, lim=31)
    g, t = w.gallery.numpy(), w.queries.numpy()
    lab = w.labels.astype(np.uint8)
    per_query = np.zeros((q, 4), dtype=np.float64)
    mean_host = ctx.eval_ranking_host(g, t, lab, c, topn, prec, per_query_out=per_query)
    dev = ctx.eval_ranking(w.gallery.cuda(), w.queries.cuda(), torch.from_numpy(lab).cuda(), c, topn, prec, 0,
                           want_per_query=True)
    assert mean_host == dev["mean"]
    assert np.array_equal(per_query, dev["per_query"].cpu().numpy())
    sample = np.arange(0, q, max(1, q // 6))
    scores = (t[sample] @ g.T).astype(np.float32)
    want, _ = spec.eval_scores_matrix(scores, w.labels, topn)
    assert np.abs(per_query[sample] - want).max() <= 1e-9
    # a second call reuses the staging buffers while the first call's kernels may still be in flight
    assert ctx.eval_ranking_host(g, t, lab, c, topn, prec) == mean_host
