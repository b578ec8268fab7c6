"""Import shim for the *unmodified* reference.  TEST / BASELINE INFRASTRUCTURE ONLY.

``/root/reference`` exists only in the build container (never on the GPU box).  There this module
is used by ``oracle/make_golden.py`` to generate ``tests/golden/*.json`` and by the
``needs_reference`` CPU tests that cross-check ``oracle/spec.py`` against the live reference.
On the GPU box the same unmodified package is found under ``baseline/_ref/`` -- installed there by
``oracle/stage_reference.py`` (``pip install --no-index --no-deps --target baseline/_ref``, run from
``__graft_entry__.build()`` in the build container; git-ignored, ships with the gpurun snapshot) --
and is what ``bench.py --impl reference`` and the ``cpu_baseline`` leg time (``kind: "reference"``).

Why a shim: the package does not import as shipped -- ``debias_clip/__init__.py:26``
does ``from measuring_bias import measure_bias`` (a top-level import) and
``debias_clip/model/model.py:4`` / ``debias_clip/datasets.py:10`` need the un-vendored
``clip`` and ``gdown`` packages.  We register empty stand-ins for those two modules and
put both ``<ref>`` and ``<ref>/debias_clip`` on ``sys.path``; nothing of the reference is
copied or modified.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _has_reference(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "debias_clip", "measuring_bias.py"))


def _find_root() -> str:
    env = os.environ.get("DVL_REFERENCE_ROOT")
    for root in ([env] if env else []) + ["/root/reference", STAGED_ROOT]:
        if _has_reference(root):
            return root
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return _has_reference(REFERENCE_ROOT)


def source() -> str:
    """Where the reference comes from: "mounted" (/root/reference), "staged" (baseline/_ref) or "absent"."""
    if not available():
        return "absent"
    return "staged" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(STAGED_ROOT) else "mounted"


def load():
    """Returns the reference's ``measuring_bias`` module (eval_ranking,
    compute_skew_metrics, normalized_discounted_KL, gen_prompts)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if "clip" not in sys.modules:
        stub = types.ModuleType("clip")
        stub.load = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("openai clip is not installed"))
        stub.tokenize = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("openai clip is not installed"))
        stub.available_models = lambda: []
        sys.modules["clip"] = stub
    if "gdown" not in sys.modules:
        stub = types.ModuleType("gdown")
        stub.download = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no network"))
        sys.modules["gdown"] = stub
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "debias_clip")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import debias_clip  # noqa: F401  (executes __init__, which imports measuring_bias top-level)
    import measuring_bias as mb
    return mb
