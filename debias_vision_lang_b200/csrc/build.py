"""Builds libdvl_b200.so (sm_100a only) in-tree with nvcc.  Usage: python build.py [--force]"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libdvl_b200.so")
SOURCES = ["api.cu", "prep.cu", "gemm.cu", "segsort.cu", "bucketsort.cu", "topk.cu", "metrics.cu"]
HEADERS = ["dvl_common.cuh", "dvl_internal.h", os.path.join("..", "..", "include", "dvl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


INFO = os.path.join(PKG, "libdvl_b200.build.json")


def sources_digest() -> str:
    import hashlib
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build_info() -> dict:
    """What the in-tree library was built from (written by build(); bench.py copies it into its JSON line).
    `sources_match` compares the digest recorded at build time with the sources present now."""
    import json
    try:
        with open(INFO) as f:
            info = json.load(f)
    except (OSError, ValueError):
        return {"built": None}
    try:
        info["sources_match"] = info.get("sources_sha16") == sources_digest()
    except OSError:
        info["sources_match"] = None
    return info


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles when the library is missing, older than a source, built from different sources, or when
    forced (argument, or DVL_FORCE_BUILD=1 in the environment)."""
    import json
    import time
    force = force or os.environ.get("DVL_FORCE_BUILD", "") == "1"
    why = "forced" if force else ("stale" if stale() else
                                  (None if build_info().get("sources_match") else "sources changed"))
    if why is None:
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc(), *NVCC_FLAGS, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc(), "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static"]
    subprocess.check_call(cmd)
    with open(INFO, "w") as f:
        json.dump({"built": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "why": why,
                   "sources_sha16": sources_digest(), "flags": " ".join(NVCC_FLAGS)}, f)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
