import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference mounted (build container only)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)["data"]


def u32_to_f32(values):
    return np.asarray(values, dtype=np.uint32).view(np.float32)


@pytest.fixture(scope="session")
def stage_cases():
    return load_golden("stage_cases.json")


@pytest.fixture(scope="session")
def ranking_cases():
    return load_golden("ranking_cases.json")


def build_ranking_workload(case):
    from debias_vision_lang_b200 import synthetic
    if case["kind"] == "dyadic_grid":
        w = synthetic.dyadic_grid(case["n"], case["d"], case["q"], case["c"], seed=case["seed"], lim=case["lim"])
    else:
        w = synthetic.adversarial(case["n"], case["d"], case["q"], case["c"], seed=case["seed"])
    assert float(w.gallery.double().sum()) == case["gallery_checksum"]
    assert int(w.labels.sum()) == case["labels_checksum"]
    topn = float(case["topn"]) if case["topn_is_float"] else int(case["topn"])
    return w, topn
