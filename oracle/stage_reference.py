"""Stages the UNMODIFIED reference package under ``baseline/_ref/`` so that it travels to the GPU box
(git-ignored, not gpurun-ignored).  TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/stage_reference.py          # no-op when /root/reference is absent or already staged

The recipe is the one the task prescribes: ``pip install --no-index --no-build-isolation --no-deps
--find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>`` (a copy because the
source tree is read-only and setuptools writes ``build/`` and ``*.egg-info`` next to ``setup.py``;
``--no-deps`` because ``ftfy``, ``clip`` and ``gdown`` are not in the offline wheelhouse -- the hot path
needs none of them, ``oracle/ref_shim.py`` stubs the two that are imported).  Nothing is edited; the
staged ``measuring_bias.py`` is compared byte for byte with the mounted one.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")
HOT = os.path.join("debias_clip", "measuring_bias.py")


def staged_ok() -> bool:
    return os.path.isfile(os.path.join(DST, HOT)) and (
        not os.path.isfile(os.path.join(SRC, HOT)) or filecmp.cmp(os.path.join(SRC, HOT), os.path.join(DST, HOT), False))


def stage(verbose: bool = False) -> str:
    """Returns "present" (already staged), "staged" (installed now) or "absent" (no reference to stage)."""
    if staged_ok():
        return "present"
    if not os.path.isfile(os.path.join(SRC, HOT)):
        return "absent"
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SRC, copy)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DST, copy]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout)
        if proc.returncode != 0:
            raise RuntimeError("pip install of the reference into baseline/_ref failed")
    if not staged_ok():
        raise RuntimeError("staged reference differs from /root/reference")
    return "staged"


if __name__ == "__main__":
    print(stage(verbose=True))
