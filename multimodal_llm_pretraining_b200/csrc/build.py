"""Builds the C-ABI CUDA libraries in-tree with nvcc for sm_100a:
    libb200pt.so       16-bit tensors are bf16 (the in-scope precision of every BASELINE.json config)
    libb200pt_fp16.so  the same sources with -DB200_ELEM_FP16: identical entry points over IEEE fp16 tensors (the reference's
                       precision for every Pythia but 1b and for RoBERTa, src/models/pythia.py:33-41)

Usage: python -m multimodal_llm_pretraining_b200.csrc.build [--force] [--verbose]
The .so files land next to the package so that they travel with the repo snapshot to the GPU box; objects go to build/
(git-ignored).
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
LIB_FP16 = PKG / "libb200pt_fp16.so"
OBJ_DIR = ROOT / "build" / "b200pt"
# variant name -> (library, object dir, extra nvcc flags)
VARIANTS = {
    "bf16": (LIB, OBJ_DIR, []),
    "fp16": (LIB_FP16, ROOT / "build" / "b200pt_fp16", ["-DB200_ELEM_FP16=1"]),
}

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


def _compile(src: Path, obj: Path, verbose: bool, extra: list[str]) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{log}")
    (obj.with_suffix(".ptxas.log")).write_text(log)
    if verbose:
        print(log)
    return log


def build(force: bool = False, verbose: bool = False) -> Path:
    """Builds every variant (stale objects only); returns the path of the bf16 library."""
    hdr_mtime = _newest_header_mtime()
    jobs = []
    link = []
    for variant, (lib, obj_dir, extra) in VARIANTS.items():
        obj_dir.mkdir(parents=True, exist_ok=True)
        objs = []
        stale_any = False
        for name in SOURCES:
            src = CSRC / name
            if not src.exists():
                raise RuntimeError(f"missing source {src}")
            obj = obj_dir / (src.stem + ".o")
            objs.append(obj)
            stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_mtime)
            if stale:
                jobs.append((src, obj, extra))
                stale_any = True
        if stale_any or not lib.exists():
            link.append((lib, objs))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile(j[0], j[1], verbose, j[2]), jobs))
    for lib, objs in link:
        # -Bsymbolic: each library binds its own (identically named) internal symbols to itself when both are loaded
        cmd = [_nvcc(), "-shared", "-o", str(lib), *map(str, objs), "-cudart", "static", "-Xlinker", "-Bsymbolic"]
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
    print(LIB_FP16)
    sys.exit(0)
