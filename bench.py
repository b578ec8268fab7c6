#!/usr/bin/env python
"""Benchmark of the bias-measurement hot path (BASELINE.json metric: gallery-query pairs/s
scored + ranked + reduced to MaxSkew/NDKL).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg5|cfg4shard]
    python bench.py --impl reference ...      # the reference's CPU path (row-by-row port) on host cores

One JSON line on stdout (rank 0).  A step = one pass of the whole hot path over one synthetic
batch: similarity -> ranking -> label histograms -> MaxSkew + NDKL -> mean over queries.
`value` = pairs (N*Q, counted once) per second with inputs resident in HBM; `e2e` = the same
through dvl_eval_ranking_host with pinned HOST buffers (H2D of gallery/queries/labels and D2H of
the result inside the timed region).  Under torchrun (N > 1) every rank holds the gallery and
ranks its own slice of the queries (top_n = N workloads shard by query: no data-path collective;
see DESIGN.md section 6), timing is max-over-ranks of CUDA-event time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "gallery-query pairs/s scored+ranked (MaxSkew/NDKL)"
UNIT = "pairs/s"


def load_traffic(workload):
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(workload)
    except (OSError, ValueError):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region: NVML polled every
    2 ms from a background thread (the timed region of the small workloads lasts tens of milliseconds,
    too short for `nvidia-smi -lms`), with `nvidia-smi` at 200 ms as the fallback."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []          # (sm MHz, reasons bitmask)
        self.stop_flag = threading.Event()
        self.sm_max = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            props = torch.cuda.get_device_properties(self.gpu)
            bus_id = f"{props.pci_domain_id:08X}:{props.pci_bus_id:02X}:{props.pci_device_id:02X}.0"
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        return pynvml, handle

    def _poll(self, pynvml, handle):
        while not self.stop_flag.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                self.samples.append((float(sm), int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            pynvml, handle = self._nvml_handle()
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, args=(pynvml, handle), daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=1.0)
            if not self.samples:
                return dict(sm_mhz=None, sm_max_mhz=self.sm_max, reasons=["no samples"])
            p = self.nvml
            names = {"hw_slowdown": p.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": p.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": p.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": p.nvmlClocksEventReasonSwPowerCap}
            seen = 0
            for _, r in self.samples:
                seen |= r
            return dict(sm_mhz=float(np.median([x for x, _ in self.samples])), sm_max_mhz=self.sm_max,
                        reasons=sorted(k for k, bit in names.items() if seen & bit), samples=len(self.samples),
                        source="nvml, 2 ms poll during the timed region")
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                    samples=len(sm), source="nvidia-smi -lms 200")


SHARDED_WORKLOADS = ("cfg3", "cfg4", "cfg5")      # top_n << N: gallery-sharded (strong scaling) at every N
FULLSORT_WORKLOADS = ("cfg1", "cfg2")             # top_n = N (measure_bias default): one GPU / query-sharded
CHECKSUMS = os.path.join(ROOT, "tests", "golden", "bench_checksums.json")


def config_of(name):
    from debias_vision_lang_b200 import synthetic
    return synthetic.BASELINE_CONFIGS[name]


def make_sharded_workload(name, device, lo, hi, n_global, q, dtype):
    """Rows [lo, hi) of the workload's gallery (+ all queries): the same rows whatever the shard layout."""
    from debias_vision_lang_b200 import synthetic
    cfg = config_of(name)
    seed = 1234 + int(name[-1])
    if name == "cfg5":           # adversarial dyadic batch: generated whole on the host (numpy), then cut
        w = synthetic.baseline_workload(name, device="cpu", n_override=n_global, q_override=q)
        w.gallery = w.gallery[lo:hi].to(device=device, dtype=dtype).contiguous()
        w.queries = w.queries.to(device=device, dtype=dtype)
        w.labels = w.labels[lo:hi]
    else:
        w = synthetic.clip_like_rows(n_global, cfg["d"], q, cfg["c"], seed, lo, hi, device=device, dtype=dtype)
    w.name, w.topn = name, cfg["topn"]
    return w


def describe_sharded(name, n_global, n_local, d, q, c, top_n, precision, input_dtype, world):
    return {
        "workload": f"{name}: {n_global}x{d} gallery x {q} queries, C={c}, top_n={top_n}, {precision} similarity"
                    + (f" (BASELINE.json configs[{int(name[-1]) - 1}])" if n_global == config_of(name)["n"] else
                       " (REDUCED gallery: development run, not the BASELINE config)"),
        "n_gallery": n_global, "depth": d, "n_queries": q, "n_classes": c, "top_n": top_n,
        "precision": precision, "input_dtype": input_dtype, "rows_per_gpu": n_local,
        "l2_policy": f"inputs larger than L2 ({n_local * d * (2 if input_dtype == 'bfloat16' else 4) / 1e6:.0f} MB "
                     "gallery shard per GPU vs 126 MB L2); no explicit flush",
        "parallelism": f"gallery-sharded x{world} (contiguous row blocks, queries replicated): local streaming top-k, "
                       "NCCL all-gather of per-shard statistics and candidate lists, deterministic merge"
                       + ("" if world > 1 else "; one shard holds the whole gallery, the collectives degenerate"),
    }


def describe_fullsort(name, n, d, q, c, world):
    return {
        "workload": f"{name}: {n}x{d} gallery x {q} queries, C={c}, top_n=N (measure_bias default), fp32 similarity"
                    + (f"; {world} GPUs, each ranks its own {q} queries over the full gallery" if world > 1 else ""),
        "n_gallery": n, "depth": d, "n_queries_per_gpu": q, "n_classes": c, "top_n": n, "precision": "fp32",
        "input_dtype": "float32",
        "l2_policy": f"inputs larger than L2 ({n * d * 4 / 1e6:.0f} MB fp32 gallery + {q * n * 4 / 1e6:.0f} MB of keys "
                     "vs 126 MB L2); no explicit flush",
        "parallelism": "1 GPU" if world == 1 else f"query-sharded x{world} (replicated gallery; one all-gather of the "
                                                   "per-query results + global mean inside the timed step)",
    }


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own eval_ranking (unmodified package staged under baseline/_ref, or
# /root/reference in the build container) on the host cores; the row-by-row port only if neither exists
# ------------------------------------------------------------------------------------------
def cpu_arm():
    """(kind, callable(labels, gallery, queries, topn) -> {"maxskew": {...}, "ndkl": {...}})."""
    from oracle import ref_shim
    if ref_shim.available():
        mb = ref_shim.load()

        def run(labels, g, t, topn):
            return {ev: mb.eval_ranking(labels, g, t, ev, topn) for ev in ("maxskew", "ndkl")}
        return "reference", run, f"unmodified debias_clip.eval_ranking ({ref_shim.source()}: {ref_shim.REFERENCE_ROOT})"
    from oracle import port
    return "port", port.measure, "row-by-row port oracle/port.py (reference package not staged)"


def cpu_workload(name, rows_cap=1_000_000):
    """The workload on the host for the CPU arm.  Galleries above `rows_cap` rows are cut to their first
    `rows_cap` rows (every per-query cost of the reference is O(N): mat-vec, Counter, nlargest -- and the
    NDKL loop runs over top_n rows either way), which keeps host memory and generation time bounded."""
    from debias_vision_lang_b200 import synthetic
    cfg = config_of(name)
    if name in SHARDED_WORKLOADS:
        n = min(cfg["n"], rows_cap)
        w = make_sharded_workload(name, "cpu", 0, n, cfg["n"] if name != "cfg5" else n, cfg["q"], torch.float32)
        if cfg["precision"] == "bf16":     # the values the GPU arm ranks: embeddings rounded to bf16
            w.gallery = w.gallery.bfloat16().float()
            w.queries = w.queries.bfloat16().float()
        cut = None if n == cfg["n"] else f"first {n} of {cfg['n']} gallery rows (per-query cost is linear in N)"
    else:
        w = synthetic.baseline_workload(name, device="cpu")
        cut = None
    return w, cut


def cpu_sample(run, w, nq):
    g, t = w.gallery, w.queries[:nq]
    t0 = time.perf_counter()
    with open(os.devnull, "w") as devnull:       # the reference prints tqdm bars
        err, sys.stderr = sys.stderr, devnull
        try:
            res = run(w.labels, g, t, w.topn)
        finally:
            sys.stderr = err
    dt = time.perf_counter() - t0
    return g.shape[0] * t.shape[0] / dt, dt, res


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t_all = time.perf_counter()
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is one process and takes every host core
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    kind, run, what = cpu_arm()
    w, cut = cpu_workload(args.workload)
    cfg = config_of(args.workload)
    _, t_one, _ = cpu_sample(run, w, 1)
    # each step = a bounded sample: K + W steps together take ~150 s of CPU time
    nq = int(150.0 / ((args.steps + args.warmup) * max(t_one, 1e-3)))
    nq = max(1, min(nq, w.queries.shape[0], 64))
    for _ in range(args.warmup):
        cpu_sample(run, w, nq)
    vals = [cpu_sample(run, w, nq)[:2] for _ in range(args.steps)]
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    cores = torch.get_num_threads()
    n, d = w.gallery.shape
    sample = (f"{nq} of {cfg['q']} queries per step" + (f", {cut}" if cut else " x full gallery")
              + f", both evaluations; {what}; torch mat-vec on {cores} threads, pandas/python part one thread")
    if args.workload in SHARDED_WORKLOADS:
        config = describe_sharded(args.workload, cfg["n"], n, d, cfg["q"], cfg["c"], cfg["topn"], cfg["precision"],
                                  "bfloat16" if cfg["precision"] == "bf16" else "float32", 1)
        config["parallelism"] = "host CPU, one process"
    else:
        config = describe_fullsort(args.workload, n, d, cfg["q"], cfg["c"], 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.workload in SHARDED_WORKLOADS else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_record(name, seconds_budget=20.0):
    """Bounded sample of the same workload through the reference on this box's host cores."""
    kind, run, what = cpu_arm()
    w, cut = cpu_workload(name)
    _, t_one, _ = cpu_sample(run, w, 1)
    nq = max(1, min(int(seconds_budget / max(t_one, 1e-3)), w.queries.shape[0], 64))
    v, dt, _ = cpu_sample(run, w, nq)
    return {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{nq} of {w.queries.shape[0]} queries" + (f", {cut}" if cut else " x full gallery")
                      + f", both evaluations ({dt:.1f} s); {what}; mat-vec multi-threaded, pandas/python part one thread",
            "host_cpus": os.cpu_count()}


# ------------------------------------------------------------------------------------------
class Harness:
    """Process-group plumbing and the timing rule: barrier + synchronize on both sides, CUDA events on
    the launching (current) stream, max over ranks."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a B200: there is no CPU fallback for the product path "
                               "(use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local_rank)
        self.dev = f"cuda:{self.local_rank}"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = torch.tensor(list(values), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def timed(self, step, steps):
        """Milliseconds for exactly `steps` calls of step(), max over ranks; returns (ms, last result)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        out = None
        for _ in range(steps):
            out = step()
        ev1.record()
        self.barrier()
        return self.max_over_ranks([ev0.elapsed_time(ev1)])[0], out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def means_of(r):
    return [r["maxskew"]["eq_opp"], r["maxskew"]["dem_par"], r["ndkl"]["eq_opp"], r["ndkl"]["dem_par"]]


def result_digest(per_query, ranked_len):
    import hashlib
    h = hashlib.sha256()
    h.update(per_query.detach().cpu().contiguous().numpy().tobytes())
    h.update(ranked_len.detach().cpu().to(torch.int32).contiguous().numpy().tobytes())
    return h.hexdigest()


def load_checksums():
    try:
        with open(CHECKSUMS) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


# ------------------------------------------------------------------------------------------
# top_n << N: gallery-sharded route (BASELINE configs 3, 4, 5), strong scaling
# ------------------------------------------------------------------------------------------
def run_sharded(h, args, name, peaks):
    from debias_vision_lang_b200 import _capi, sharded
    cfg = config_of(name)
    n_global = args.rows or cfg["n"]
    q = args.queries or cfg["q"]
    precision = cfg["precision"]
    in_dtype = torch.bfloat16 if (precision == "bf16" and args.input_dtype != "fp32") else torch.float32
    lo, hi = sharded.shard_bounds(n_global, h.world, h.rank)
    w = make_sharded_workload(name, h.dev, lo, hi, n_global, q, in_dtype)
    n_local, d = w.gallery.shape
    top_n = int(w.topn)
    lab = torch.from_numpy(w.labels.astype(np.uint8)).to(h.dev)
    ctx = _capi.context(h.local_rank)
    comm = sharded.TorchDistComm() if h.world > 1 else sharded.SoloComm()
    opts = {"precision": precision}
    g_local, queries = w.gallery, w.queries

    def step():
        return sharded.evaluate_sharded(ctx, comm, g_local, queries, lab, lo, n_global, w.n_classes, top_n, opts)

    for _ in range(max(args.warmup, 3)):
        r = step()
    h.barrier()
    warm_ms, r = h.timed(step, 1)
    # Per-stage device times come from event pairs around every stage (~40 records per step, no graph on this
    # route).  Next to a 10-100 ms step that is nothing and the stages are timed INSIDE the timed region; a
    # sub-millisecond step (config 3 / 5 shards) paid ~10 % for them (0.52 vs 0.45 ms on a 125 k-row shard), so
    # there the timed loop runs bare and the stages are timed over a second loop right after it.
    profile_inline = warm_ms >= 5.0
    ctx.get_profile()
    ctx.set_profiling(profile_inline)
    launches0 = ctx.launch_count
    sampler = ClockSampler(h.local_rank)
    sampler.start()
    elapsed_ms, r = h.timed(step, args.steps)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    profiled_steps, profiled_ms = args.steps, elapsed_ms
    if not profile_inline:
        profiled_steps = max(3, min(args.steps, 20))
        ctx.get_profile()
        ctx.set_profiling(True)
        profiled_ms, _ = h.timed(step, profiled_steps)
    profile = ctx.get_profile()
    ctx.set_profiling(False)
    pairs_per_step = n_global * q
    value = pairs_per_step * args.steps / (elapsed_ms * 1e-3)
    names = sorted(profile)
    stage_ms = {k: profile[k][0] / profiled_steps for k in names}
    stage_launches = {k: profile[k][1] // profiled_steps for k in names}
    stage_ms_max = dict(zip(names, h.max_over_ranks([stage_ms[k] for k in names]))) if h.world > 1 else None

    # ---- self-check: per-query results against the committed one-GPU digest of the same workload ----
    key = f"{name}:{n_global}x{d}x{q}:top{top_n}:{precision}:{str(in_dtype).replace('torch.', '')}"
    digest = result_digest(r["per_query"], r["ranked_len"])
    expected = load_checksums().get(key)
    verify = {"key": key, "per_query_sha256": digest, "expected_from_1gpu": expected,
              "match": (digest == expected) if expected else None, "n_flagged": int(r["n_flagged"]),
              "list_len": int(r["list_len"])}

    # ---- end to end: the local shard starts in pinned HOST memory every step ----
    def host_copy(t):
        try:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        except RuntimeError:          # the box cannot pin this much: pageable memory (slower, still correct)
            buf = torch.empty(t.shape, dtype=t.dtype)
        buf.copy_(t)
        return buf

    e2e_steps = max(3, min(args.steps, 10))
    host_ok = 1.0
    try:
        g_host, l_host, t_host = host_copy(g_local), host_copy(lab), host_copy(queries)
    except (RuntimeError, MemoryError):       # a host without room for a copy of the shard: no end-to-end number
        host_ok = 0.0
    torch.cuda.synchronize()
    if h.max_over_ranks([-host_ok])[0] == -1.0:       # every rank has its copy
        def e2e_step():
            return sharded.evaluate_sharded_host(ctx, comm, g_host, t_host, l_host, lo, n_global, w.n_classes, top_n, opts)

        for _ in range(2):
            re = e2e_step()
        e2e_ms, re = h.timed(e2e_step, e2e_steps)
        assert means_of(re) == means_of(r), "host-input and device-input entry points disagree"
        h2d = (g_host.numel() * g_host.element_size() + l_host.numel() + t_host.numel() * t_host.element_size())
        h2d_all = int(h2d) * h.world       # this rank's bytes x ranks (the shards differ by < 256 rows)
        e2e = {"value": pairs_per_step * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / e2e_steps,
               "steps": e2e_steps, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": 56 * h.world,
               "pinned": bool(g_host.is_pinned()),
               "api": "sharded.evaluate_sharded_host -> dvl_shard_topk_host (C ABI, pinned host gallery shard/queries/"
                      "labels; row-chunked H2D on a copy stream overlapped with the streaming top-k), then the exchange"}
        del g_host
    else:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "unavailable": "a rank could not allocate a host copy of its gallery shard"}

    # ---- roofline: the similarity GEMM (the one dense contraction), tensor-bound ----
    flops = 2.0 * n_local * q * d                      # algorithmic 2*D flop per pair, this rank's pairs
    terms = 3 if precision == "fp32" else 1
    gemm_ms = stage_ms["gemm"]
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "sim_gemm_kernel (FILTER epilogue; includes the pilot/seed launches)",
            "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"], "traffic": None,
            "peak_source": peaks["source"] + " bf16 burst (cuBLAS 8192^3 best of 10)",
            "peak_sustained": peaks["bf16_tflops_sustained"],
            "frac_of_sustained": achieved * terms / peaks["bf16_tflops_sustained"],
            "issued_tflops": achieved * terms, "ms_per_step": gemm_ms,
            "launches_per_step": stage_launches.get("gemm"),
            "algorithmic_flop_per_step_per_gpu": flops,
            "share_of_step": gemm_ms / (profiled_ms / profiled_steps),
            "note": f"algorithmic 2*D flop/pair x {n_local}x{q} pairs on rank 0; fp32 mode issues {terms}x that in "
                    "16-bit tensor-core products; long steps run at power-capped clocks, hence frac_of_sustained"}
    ncu = load_traffic(name)
    if ncu is not None:
        roof["traffic"] = ncu.get("traffic_bytes_per_launch")
        roof["traffic_source"] = ncu.get("source")
        if "tensor_pipe_active_pct" in ncu:
            roof["ncu_tensor_pipe_active_pct"] = ncu["tensor_pipe_active_pct"]

    return {
        "value": value, "ms_per_step": elapsed_ms / args.steps, "scaling": "strong",
        "dtype": "f32" if precision == "fp32" else "bf16",
        "config": describe_sharded(name, n_global, n_local, d, q, w.n_classes, top_n, precision,
                                   str(in_dtype).replace("torch.", ""), h.world),
        "roofline": roof, "stage_ms": stage_ms, "stage_launches": stage_launches,
        "stage_timing": ("inside the timed region" if profile_inline else
                         f"second loop of {profiled_steps} steps with stage events, {profiled_ms / profiled_steps:.4f} ms/step "
                         "(the timed loop of a sub-millisecond step runs without them)"),
        "stage_ms_max_over_ranks": stage_ms_max, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "verify": verify, "result": dict(zip(("maxskew_eq_opp", "maxskew_dem_par", "ndkl_eq_opp", "ndkl_dem_par"),
                                             means_of(r))),
    }


# ------------------------------------------------------------------------------------------
# top_n = N: full-sort route (BASELINE configs 1, 2 -- what measure_bias runs by default)
# ------------------------------------------------------------------------------------------
def run_fullsort(h, args, name, peaks, steps, warmup):
    from debias_vision_lang_b200 import _capi, measuring_bias, sharded, synthetic
    cfg = config_of(name)
    w = synthetic.baseline_workload(name, device=h.dev)
    world = h.world
    if world > 1:     # query sharding: same gallery on every rank, a different slice of a world-times larger query set
        w.queries = synthetic.clip_like(1024, cfg["d"], cfg["q"], cfg["c"], seed=4321 + h.rank, device=h.dev).queries
    n, d = w.gallery.shape
    q = w.queries.shape[0]
    top_n = n if w.topn == 1.0 else int(w.topn)
    prec = _capi.PRECISION_FP32
    lab = torch.from_numpy(w.labels.astype(np.uint8)).to(h.dev)
    ctx = _capi.context(h.local_rank)
    if world > 1:
        comm = sharded.TorchDistComm()
        all_queries = comm.all_gather(w.queries.contiguous()).reshape(world * q, d).contiguous()
        local_reference = ctx.eval_ranking(w.gallery, w.queries, lab, w.n_classes, top_n, prec)["mean"]

        def step():
            return means_of(sharded.evaluate_query_sharded(
                ctx, comm, w.gallery, all_queries, lab, w.n_classes, top_n,
                {"precision": "fp32", "ranked_len": False, "per_query": False}))
    else:
        def step():
            return ctx.eval_ranking(w.gallery, w.queries, lab, w.n_classes, top_n, prec)["mean"]

    # the product path: stream-ordered launches, replayed as a CUDA graph from the third identical call on
    for _ in range(max(warmup, 3)):
        result = step()
    launches0 = ctx.launch_count
    sampler = ClockSampler(h.local_rank)
    sampler.start()
    elapsed_ms, result = h.timed(step, steps)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    # the same steps again with per-stage events (direct launches: events cannot bracket stages of a graph)
    ctx.get_profile()
    ctx.set_profiling(True)
    prof_ms, _ = h.timed(step, steps)
    profile = ctx.get_profile()
    ctx.set_profiling(False)
    pairs_per_step = n * q * world
    pairs_rank = n * q
    names = sorted(profile)
    stage_ms = {k: profile[k][0] / steps for k in names}
    stage_launches = {k: profile[k][1] // steps for k in names}

    # ---- roofline per SURVEY 8(d): unfused k = N ranking = Q*N*4 B of scores in + Q*N*4 B of order out ----
    rank_ms = stage_ms["sort_hist"] + stage_ms["sort_scan"] + stage_ms["sort_scatter"]
    alg_bytes = 8.0 * pairs_rank
    achieved = alg_bytes / (rank_ms * 1e-3) / 1e9
    ncu = load_traffic(name) or {}
    roof = {"bound": "hbm", "kernel": ncu.get("kernels", "ranking stage (segmented sort kernels)"),
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": ncu.get("traffic_bytes_per_step"), "traffic_source": ncu.get("source"),
            "traffic_ratio": (ncu["traffic_bytes_per_step"] / alg_bytes) if ncu.get("traffic_bytes_per_step") else None,
            "peak_source": peaks["source"] + " copy bandwidth",
            "algorithmic_bytes_per_step": alg_bytes, "ms_per_step": rank_ms,
            "launches_per_step": sum(stage_launches[k] for k in ("sort_hist", "sort_scan", "sort_scatter")),
            "share_of_step": rank_ms / (prof_ms / steps),
            "note": "algorithmic 8 B per pair per STEP (SURVEY 8d) over the time of every ranking-stage kernel of the step"}

    # ---- end to end through the C ABI with pinned host buffers ----
    g_host = w.gallery.detach().float().cpu().pin_memory()
    t_host = w.queries.detach().float().cpu().pin_memory()
    l_host = torch.from_numpy(w.labels.astype(np.uint8)).pin_memory()
    g_np, t_np, l_np = g_host.numpy(), t_host.numpy(), l_host.numpy()
    for _ in range(2):
        e2e_result = ctx.eval_ranking_host(g_np, t_np, l_np, w.n_classes, top_n, prec)
    e2e_ms, e2e_result = h.timed(lambda: ctx.eval_ranking_host(g_np, t_np, l_np, w.n_classes, top_n, prec), steps)
    device_result = result if world == 1 else local_reference      # at N > 1 `result` is the mean over all ranks
    assert max(abs(a - b) for a, b in zip(e2e_result, device_result)) == 0.0, "host and device entry points disagree"
    h2d_bytes = int(g_np.nbytes + t_np.nbytes + l_np.nbytes) * world
    e2e = {"value": pairs_per_step * steps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / steps,
           "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 40 * world,
           "api": "dvl_eval_ranking_host (C ABI, pinned host buffers)"}

    # ---- the same call with the gallery held in fp16 on the host (CLIP on CUDA emits fp16 embeddings):
    # half the PCIe bytes.  The fp16 rows are a (slightly) different gallery, so the check is against the device
    # entry point on the same fp16 rows.
    e2e_f16 = None
    if world == 1:
        g16_host = w.gallery.detach().half().cpu().pin_memory()
        want16 = ctx.eval_ranking(w.gallery.half(), w.queries, lab, w.n_classes, top_n, prec)["mean"]
        for _ in range(2):
            got16 = ctx.eval_ranking_host(g16_host, t_host, l_host, w.n_classes, top_n, prec)
        f16_ms, got16 = h.timed(lambda: ctx.eval_ranking_host(g16_host, t_host, l_host, w.n_classes, top_n, prec), steps)
        assert got16 == want16, "fp16 host gallery: host and device entry points disagree"
        e2e_f16 = {"value": pairs_per_step * steps / (f16_ms * 1e-3), "unit": UNIT, "ms_per_step": f16_ms / steps,
                   "h2d_bytes_per_step": int(g16_host.numel() * 2 + t_np.nbytes + l_np.nbytes),
                   "note": "gallery rows as fp16 in pinned host memory (what CLIP on CUDA produces), fp32 mode arithmetic"}
        del g16_host

    # ---- the call a user makes: the Python drop-in with int64 numpy labels and host (pinned) tensors ----
    labels_np = w.labels.astype(np.int64)

    def api_step():
        return means_of(measuring_bias.evaluate(labels_np, g_host, t_host, 1.0, {"precision": "fp32"}))

    for _ in range(2):
        api_result = api_step()
    api_ms, api_result = h.timed(api_step, steps)
    assert max(abs(a - b) for a, b in zip(api_result, device_result)) == 0.0, "Python drop-in and C ABI disagree"
    api_e2e = {"value": pairs_per_step * steps / (api_ms * 1e-3), "unit": UNIT, "ms_per_step": api_ms / steps,
               "api": "debias_vision_lang_b200.measuring_bias.evaluate(labels int64 ndarray, pinned host tensors) -- what "
                      "eval_ranking / measure_bias call", "vs_c_abi_e2e": (api_ms / e2e_ms)}
    # ... and with the embeddings already on the device (the encoders' output), labels still an int64 ndarray
    api_resident = None
    if world == 1:
        def api_dev_step():
            return means_of(measuring_bias.evaluate(labels_np, w.gallery, w.queries, 1.0, {"precision": "fp32"}))
        for _ in range(3):
            api_dev = api_dev_step()
        apid_ms, api_dev = h.timed(api_dev_step, steps)
        assert max(abs(a - b) for a, b in zip(api_dev, device_result)) == 0.0, "Python drop-in (device) and C ABI disagree"
        api_resident = {"value": pairs_per_step * steps / (apid_ms * 1e-3), "unit": UNIT, "ms_per_step": apid_ms / steps,
                        "api": "measuring_bias.evaluate(labels int64 ndarray, CUDA tensors): label upload + "
                               "dvl_encode_labels + dvl_eval_ranking", "vs_c_abi_resident": apid_ms / elapsed_ms}
    return {
        "value": pairs_per_step * steps / (elapsed_ms * 1e-3), "ms_per_step": elapsed_ms / steps, "scaling": "weak",
        "dtype": "f32", "config": describe_fullsort(name, n, d, q, w.n_classes, world), "roofline": roof,
        "profiled_ms_per_step": prof_ms / steps, "stage_ms": stage_ms, "stage_launches": stage_launches,
        "e2e": e2e, "e2e_f16_host": e2e_f16, "api_e2e": api_e2e, "api_resident": api_resident, "gpu_launches": int(launches), "clocks": clocks,
        "launch_mode": "CUDA-graph replay of the route (timed loop); direct launches in the profiled loop",
        "result": dict(zip(("maxskew_eq_opp", "maxskew_dem_par", "ndkl_eq_opp", "ndkl_dem_par"), result)),
    }


def run_ours(args):
    from debias_vision_lang_b200._capi import build_info
    h = Harness()
    peaks = load_peaks()
    name = args.workload
    if name in SHARDED_WORKLOADS:
        rec = run_sharded(h, args, name, peaks)
        # the k = N regime (measure_bias's default call) stays tracked beside it: one GPU, rank 0's record
        secondary = None
        if h.world == 1 and not args.no_secondary:
            torch.cuda.empty_cache()
            secondary = run_fullsort(h, args, "cfg2", peaks, max(args.steps, 20), max(args.warmup, 5))
            secondary.pop("clocks", None)
    else:
        rec = run_fullsort(h, args, name, peaks, args.steps, args.warmup)
        secondary = None
    cpu = None
    if h.rank == 0 and h.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_record(name)
    if h.rank == 0:
        line = {"metric": METRIC, "value": rec.pop("value"), "unit": UNIT, "n_gpus": h.world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True,
                "scaling": rec.pop("scaling"), "vs_baseline": None, "dtype": rec.pop("dtype"),
                "data": "synthetic (seeded CLIP-like embeddings and labels; no FairFace/CLIP weights offline)",
                "config": rec.pop("config")}
        line.update(rec)
        line["cpu_baseline"] = cpu
        if secondary is not None:
            line["secondary"] = secondary
        line["build"] = build_info()
        print(json.dumps(line), flush=True)
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json config; default cfg4 = the configuration the metric is quoted on "
                         "(10 M x 768 gallery, 10 k queries, top-1000, bf16), gallery-sharded over --gpus")
    ap.add_argument("--rows", type=int, default=0, help="development: cut the gallery of a sharded workload to this many rows")
    ap.add_argument("--queries", type=int, default=0, help="development: number of queries of a sharded workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg2 (top_n = N) sub-record")
    ap.add_argument("--input-dtype", default="auto", choices=["auto", "fp32", "bf16"],
                    help="dtype of the resident embeddings of a bf16-similarity workload: auto/bf16 = bf16 rows read in "
                         "place by TMA; fp32 = fp32 rows converted by the library's split pass every step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
