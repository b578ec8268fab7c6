"""bench.py's reference arm on a tiny workload (CPU): the JSON line carries every key the driver reads,
the arm runs the reference (or, where it is not staged, the port) through its own eval_ranking, and a rank
other than 0 prints nothing.  The GPU arm's line is exercised on the GPU box by the driver itself."""
import argparse
import io
import json
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from debias_vision_lang_b200 import synthetic  # noqa: E402

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _tiny(name, rows_cap=0):
    w = synthetic.clip_like(600, 32, 5, 2, seed=3)
    w.name, w.topn = name, 50
    return w, "tiny test workload"


def _args():
    return argparse.Namespace(workload="cfg4", steps=1, warmup=0, gpus=1, impl="reference")


def test_reference_arm_line(monkeypatch):
    monkeypatch.setattr(bench, "cpu_workload", _tiny)
    monkeypatch.delenv("RANK", raising=False)
    out = io.StringIO()
    with redirect_stdout(out):
        bench.run_reference(_args())
    lines = [l for l in out.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert REQUIRED <= set(line)
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["higher_is_better"] is True
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_is_silent_off_rank_zero(monkeypatch):
    monkeypatch.setattr(bench, "cpu_workload", _tiny)
    monkeypatch.setenv("RANK", "1")
    out = io.StringIO()
    with redirect_stdout(out):
        bench.run_reference(_args())
    assert out.getvalue() == ""
