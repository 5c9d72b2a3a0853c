"""Builds libb200pt.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python -m multimodal_llm_pretraining_b200.csrc.build [--force] [--verbose]
The .so lands next to the package (multimodal_llm_pretraining_b200/libb200pt.so) so that it travels with the repo
snapshot to the GPU box; objects go to build/ (git-ignored).
"""

from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
ROOT = PKG.parent
LIB = PKG / "libb200pt.so"
OBJ_DIR = ROOT / "build" / "b200pt"

SOURCES = [
    "api.cu",
    "layernorm.cu",
    "elementwise.cu",
    "cross_entropy.cu",
    "adam.cu",
    "gemm.cu",
    "attention.cu",
]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas",
    "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libb200pt cannot be built")


def _newest_header_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "b200pt.h"]
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: Path, obj: Path, verbose: bool) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{log}")
    (obj.with_suffix(".ptxas.log")).write_text(log)
    if verbose:
        print(log)
    return log


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    hdr_mtime = _newest_header_mtime()
    jobs = []
    objs = []
    for name in SOURCES:
        src = CSRC / name
        if not src.exists():
            raise RuntimeError(f"missing source {src}")
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_mtime)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(j[0], j[1], verbose), jobs))
    if jobs or not LIB.exists():
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}{res.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    path = build(force=a.force, verbose=a.verbose)
    print(path)
    sys.exit(0)
