"""Builds libvfidkr_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m vfidkr_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the tree.
Only sm_100a SASS is emitted (no PTX for other targets, no multi-arch fatbin).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
OBJ_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "libvfidkr_b200.so"

SOURCES = ["capi.cu", "filterinterpolation.cu", "fi_strip.cu", "fi_strip_w128.cu", "fi_strip_dkr.cu", "fi_bigc.cu", "projection.cu", "interpolation.cu", "separableconv.cu",
           "correlation.cu", "correlation_tc.cu", "pwcwarp.cu", "frameio.cu"]

# sources that #include another source (same kernel, other compile-time geometry)
INCLUDES_SOURCE = {"fi_strip_w128.cu": ["fi_strip.cu"]}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-I", str(INCLUDE), "-I", str(CSRC),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def _host_compiler_flags() -> list[str]:
    # the image exports CC/CXX pointing at a wrapper; nvcc is happiest with the system g++
    if Path("/usr/bin/g++").exists():
        return ["-ccbin", "/usr/bin/g++"]
    return []


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """VFIDKR_NVCC_EXTRA (space-separated) adds flags for debug builds, e.g. -DVFIDKR_STRIP_STATS; use with --force."""
    nvcc = _nvcc()
    extra = os.environ.get("VFIDKR_NVCC_EXTRA", "").split()
    OBJ_DIR.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h")) + [Path(__file__)]
    jobs = []
    for name in SOURCES:
        src, obj = CSRC / name, OBJ_DIR / (name + ".o")
        if force or _stale(obj, [src] + [CSRC / d for d in INCLUDES_SOURCE.get(name, [])] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + _host_compiler_flags() + NVCC_FLAGS + extra + ["-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for msg in ex.map(compile_one, jobs):
                if verbose and msg.strip():
                    print(msg)
    objs = [OBJ_DIR / (name + ".o") for name in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [nvcc] + _host_compiler_flags() + ["-shared", "-cudart", "static", "-gencode",
                                                  "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH)] + [str(o) for o in objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(p)
