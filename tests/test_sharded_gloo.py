"""The gallery-sharded protocol (debias_vision_lang_b200.sharded.evaluate_sharded) under a real
process group: world_size 2 (and 3) gloo on CPU, with the C-ABI stages replaced by an oracle-backed
fake so that no GPU is needed.  Checks the collective sequence, shard bounds, cross-shard exact
ties, empty shards and the flagged-query re-rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from debias_vision_lang_b200 import sharded, synthetic
from oracle import spec


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_ctx import FakeCtx
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synthetic.adversarial(case["n"], 32, case["q"], case["c"], seed=case["seed"], n_shards=world)
        lo, hi = sharded.shard_bounds(case["n"], world, rank)
        comm = sharded.TorchDistComm()
        if case.get("route") == "query":      # top_n = N: replicated gallery, query slices
            res = sharded.evaluate_query_sharded(FakeCtx(), comm, w.gallery, w.queries,
                                                 torch.from_numpy(w.labels.astype(np.uint8)), case["c"], case["topn"],
                                                 ranked_cap=case["topn"] + 8)
            torch.save({k: v for k, v in res.items()}, os.path.join(out_dir, f"rank{rank}.pt"))
            return
        if "spec" in case:      # force the speculative list length to be too short: second, exact exchange
            sharded.speculative_list_len = lambda *a, **k: case["spec"]
        res = sharded.evaluate_sharded(FakeCtx(list_cap=case["cap"]), comm, w.gallery[lo:hi].contiguous(), w.queries,
                                       torch.from_numpy(w.labels[lo:hi].astype(np.uint8)), lo, case["n"], case["c"],
                                       case["topn"], ranked_cap=case["topn"] + 8)
        torch.save({k: v for k, v in res.items()}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [
    (2, dict(n=1500, q=6, c=2, topn=40, seed=5, cap=4096)),
    (2, dict(n=700, q=5, c=3, topn=25, seed=6, cap=48)),       # small lists: tie-overflow -> flagged re-rank
    (3, dict(n=300, q=4, c=2, topn=10, seed=7, cap=4096)),      # 256-row blocks: rank 2 owns no rows
    (2, dict(n=1500, q=6, c=2, topn=40, seed=8, cap=4096, spec=8)),   # guess too short -> re-exchange
    (2, dict(n=400, q=5, c=3, topn=400, seed=9, cap=4096, route="query")),   # whole gallery ranked, query slices
    (3, dict(n=300, q=2, c=2, topn=300, seed=10, cap=4096, route="query")),  # a rank without queries
    (2, dict(n=8600, q=2, c=2, topn=8300, seed=12, cap=4096)),   # top_n beyond the streaming lists: dense route
])
def test_sharded_protocol_gloo(tmp_path, world, case):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    w = synthetic.adversarial(case["n"], 32, case["q"], case["c"], seed=case["seed"], n_shards=world)
    want = spec.measure(w.labels, w.gallery, w.queries, case["topn"])
    scores = np.stack([spec.similarity_f32(w.gallery, w.queries[i]) for i in range(case["q"])])
    _, orders = spec.eval_scores_matrix(scores, w.labels, case["topn"])
    for r, res in enumerate(outs):
        for ev in ("maxskew", "ndkl"):
            for k in ("eq_opp", "dem_par"):
                assert abs(res[ev][k] - want[ev][k]) <= 1e-12, (r, ev, k)
        for i, o in enumerate(orders):
            m = int(res["ranked_len"][i])
            assert m == len(o)
            cap = res["ranked_idx"].shape[1]
            assert res["ranked_idx"][i, :min(m, cap)].tolist() == o[:cap].tolist()
        # every rank ends with the same numbers
        assert torch.equal(res["per_query"], outs[0]["per_query"])
    if case["cap"] < 100:
        assert outs[0]["n_flagged"] > 0
    if "spec" in case:
        assert all(o["list_len"] > case["spec"] for o in outs)


def test_shard_bounds_cover_the_gallery():
    for n in (1, 255, 256, 257, 10954, 108501, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                lo, hi = sharded.shard_bounds(n, world, r)
                assert lo == covered or (lo == n and hi == n)
                assert lo % 256 == 0 or lo == n
                covered = max(covered, hi)
            assert covered == n
