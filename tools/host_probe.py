#!/usr/bin/env python
"""Developer tool: where the HOST spends a step of the gallery-sharded route (one GPU, SoloComm).

    python tools/host_probe.py [--workload cfg3] [--rows 125000] [--steps 200]

Wraps the context's stage calls with time.perf_counter (no device synchronisation added) and prints, per
call, the host microseconds inside it and between it and the previous call, averaged over back-to-back
steps -- next to the step's wall time.  The difference between the wall time and the device time of the
kernels (bench.py's stage_ms) is what the host adds: launch latency at the start of a step, the wake-up
after the one synchronisation, and result handling."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--rows", type=int, default=125000)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    import bench
    from debias_vision_lang_b200 import _capi, sharded
    dev = "cuda:0"
    torch.cuda.set_device(0)
    cfg = bench.config_of(args.workload)
    precision = cfg["precision"]
    n = args.rows or cfg["n"]
    in_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    w = bench.make_sharded_workload(args.workload, dev, 0, n, n, cfg["q"], in_dtype)
    lab = torch.from_numpy(w.labels.astype(np.uint8)).to(dev)
    ctx = _capi.context(0)
    comm = sharded.SoloComm()
    marks = []

    def wrap(obj, name):
        fn = getattr(obj, name)

        def inner(*a, **k):
            t0 = time.perf_counter()
            out = fn(*a, **k)
            marks.append((name, t0, time.perf_counter()))
            return out
        setattr(obj, name, inner)

    for name in ("shard_topk", "shard_stats", "shard_reduce", "shard_trim_pack", "merge_topk_into", "shard_finish"):
        wrap(ctx, name)
    top_n = int(w.topn)
    opts = {"precision": precision}

    def step():
        t0 = time.perf_counter()
        marks.append(("step_begin", t0, t0))
        sharded.evaluate_sharded(ctx, comm, w.gallery, w.queries, lab, 0, n, w.n_classes, top_n, opts)
        t1 = time.perf_counter()
        marks.append(("step_end", t1, t1))

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    marks.clear()
    t_all = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t_all) / args.steps
    inside, before, order = {}, {}, []
    prev_end = None
    for name, t0, t1 in marks:
        if name not in inside:
            inside[name], before[name] = 0.0, 0.0
            order.append(name)
        inside[name] += t1 - t0
        if prev_end is not None and name != "step_begin":
            before[name] += t0 - prev_end
        prev_end = t1
    print(f"wall per step {wall * 1e6:8.1f} us   ({n} x {cfg['q']} rows x queries, top_n {top_n})")
    for name in order:
        print(f"{name:18s} inside {inside[name] / args.steps * 1e6:8.1f} us   host gap before {before[name] / args.steps * 1e6:8.1f} us")


if __name__ == "__main__":
    main()
