#!/usr/bin/env python
"""Per-phase device timing of the gallery-sharded protocol (sharded.evaluate_sharded), every rank.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_trace.py [--workload cfg4]

Wraps the context's stage calls and the communicator's collectives with CUDA events (no change to the
product code) and prints, per rank, the milliseconds between consecutive phase ends of one step, averaged
over the timed steps.  "wait" inside a collective's time includes waiting for the slowest rank.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--rows", type=int, default=0, help="cut the gallery to this many rows (development)")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import bench
    from debias_vision_lang_b200 import _capi, sharded
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    cfg = bench.config_of(args.workload)
    precision = cfg["precision"]
    n_global = args.rows or cfg["n"]
    lo, hi = sharded.shard_bounds(n_global, world, rank)
    in_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    w = bench.make_sharded_workload(args.workload, dev, lo, hi, n_global, cfg["q"], in_dtype)
    g_local, index_base = w.gallery, lo
    l_local = torch.from_numpy(w.labels.astype(np.uint8)).to(dev)
    ctx = _capi.context(local)
    comm = sharded.TorchDistComm()
    marks = []

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def inner(*a, **k):
            out = fn(*a, **k)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((label, ev))
            return out
        setattr(obj, name, inner)

    for name in ("shard_topk", "shard_stats", "shard_reduce", "shard_trim_pack", "merge_topk_into", "shard_finish"):
        wrap(ctx, name, name)
    wrap(comm, "all_gather_into", "all_gather")
    top_n = int(w.topn)
    opts = {"precision": precision}

    def step():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(("start", ev))
        sharded.evaluate_sharded(ctx, comm, g_local, w.queries, l_local, index_base, n_global, w.n_classes, top_n, opts)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    marks.clear()
    acc = {}
    order = []
    for _ in range(args.steps):
        dist.barrier()
        torch.cuda.synchronize()
        step()
        torch.cuda.synchronize()
        seen = {}
        for (l0, e0), (l1, e1) in zip(marks[:-1], marks[1:]):
            seen[l1] = seen.get(l1, 0) + 1
            key = f"{l1}#{seen[l1]}"
            if key not in acc:
                acc[key] = 0.0
                order.append(key)
            acc[key] += e0.elapsed_time(e1)
        marks.clear()
    mine = {k: acc[k] / args.steps for k in order}
    mine["total"] = sum(mine.values())
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        keys = order + ["total"]
        print(json.dumps({"workload": args.workload, "world": world, "phases": keys,
                          "ms_per_rank": [[round(r[k], 4) for k in keys] for r in allr]}))
        for k in keys:
            vals = [r[k] for r in allr]
            print(f"{k:16s} min {min(vals):8.3f}  max {max(vals):8.3f}  ms")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
