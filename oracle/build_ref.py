#!/usr/bin/env python
"""Builds the reference's OWN CUDA extensions, unmodified, for sm_100a -> oracle/_ref/*.so

TEST INFRASTRUCTURE ONLY.  Nothing in the product package, bench.py's own arm or the C ABI touches these
modules; they exist to pin the CPU oracle (oracle/vfidkr_oracle.c) against the real reference kernels:
oracle/make_golden.py runs them on a B200 (through gpurun) and writes the fixtures under tests/golden/.

The sources are compiled WHERE THEY LIE under /root/reference (read-only); no reference source is copied
into this repository and the only outputs are the Python extension modules in oracle/_ref/ (git-ignored,
not gpurun-ignored, so they travel to the GPU box).  The reference's own build system (setup.py +
compiler_args.py: -std=c++11, sm_37..sm_75, torch 1.4) is NOT run; this is the short recipe instead:

    nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -include oracle/ref_shim.h  <name>_cuda_kernel.cu
    g++  -std=c++17                                -include oracle/ref_shim.h             <name>_cuda.cc
    g++  -shared ... -ltorch -ltorch_python -lc10 -lc10_cuda -lcudart

The one thing torch 2.11 no longer offers that the sources use -- AT_DISPATCH_FLOATING_TYPES on
`tensor.type()` -- is restored by the force-included overload in oracle/ref_shim.h.

    python oracle/build_ref.py [--force] [name ...]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("VFIDKR_REFERENCE", "/root/reference"))
SHIM = HERE / "ref_shim.h"

# module name -> directory holding <name>.cc, <name>_kernel.cu
EXTENSIONS = {
    "filterinterpolation_cuda": "my_package/FilterInterpolation",
    "flowprojection_cuda": "my_package/FlowProjection",
    "depthflowprojection_cuda": "my_package/DepthFlowProjection",
    "interpolation_cuda": "my_package/Interpolation",
    "interpolationch_cuda": "my_package/InterpolationCh",
    "separableconv_cuda": "my_package/SeparableConv",
    "separableconvflow_cuda": "my_package/SeparableConvFlow",
    "correlation_cuda": "PWCNet/correlation_package_pytorch1_0",
}


def available() -> bool:
    return REF.is_dir() and (REF / "my_package").is_dir()


def module_path(name: str) -> Path:
    return OUT / f"{name}.so"


def _flags(name: str):
    import torch
    tdir = Path(torch.__file__).resolve().parent
    inc = [f"-I{tdir / 'include'}", f"-I{tdir / 'include' / 'torch' / 'csrc' / 'api' / 'include'}",
           f"-I{sysconfig.get_paths()['include']}", "-I/usr/local/cuda/include"]
    defs = [f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    return tdir, inc, defs


def build_one(name: str, force: bool = False) -> Path:
    src_dir = REF / EXTENSIONS[name]
    cc, cu = src_dir / f"{name}.cc", src_dir / f"{name}_kernel.cu"
    target = module_path(name)
    if not force and target.exists() and target.stat().st_mtime >= max(cc.stat().st_mtime, cu.stat().st_mtime,
                                                                       SHIM.stat().st_mtime):
        return target
    tdir, inc, defs = _flags(name)
    OUT.mkdir(exist_ok=True)
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    with tempfile.TemporaryDirectory(prefix=f"vfidkr_ref_{name}_") as tmp:
        o_cu, o_cc = Path(tmp) / "kernel.o", Path(tmp) / "glue.o"
        cmds = [
            ["nvcc", "-ccbin", gxx, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-w",
             "-Xcompiler", "-fPIC", "-include", str(SHIM)] + inc + defs + ["-c", str(cu), "-o", str(o_cu)],
            [gxx, "-O2", "-std=c++17", "-w", "-fPIC", "-include", str(SHIM)] + inc + defs + ["-c", str(cc), "-o", str(o_cc)],
            [gxx, "-shared", "-o", str(target), str(o_cu), str(o_cc), f"-L{tdir / 'lib'}", "-L/usr/local/cuda/lib64",
             "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
             f"-Wl,-rpath,{tdir / 'lib'}"],
        ]
        for cmd in cmds:
            r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(src_dir))
            if r.returncode != 0:
                raise RuntimeError(f"{name}: {' '.join(cmd[:3])} ... failed\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
    return target


def build_all(names=None, force: bool = False, verbose: bool = False) -> dict:
    """Returns {name: path} for every extension that built.  Raises if the reference tree is absent."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    names = list(names or EXTENSIONS)
    with ThreadPoolExecutor(max_workers=min(len(names), os.cpu_count() or 4)) as ex:
        paths = list(ex.map(lambda n: build_one(n, force), names))
    if verbose:
        for n, p in zip(names, paths):
            print(f"{n}: {p}")
    return dict(zip(names, paths))


def load(name: str):
    """Import a built reference extension (torch must be importable; CUDA needed to call into it)."""
    import importlib.util

    import torch  # noqa: F401  (loads libtorch before the extension)
    p = module_path(name)
    if not p.exists():
        raise ImportError(f"{p} is missing: run `python oracle/build_ref.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    build_all(args or None, force="--force" in sys.argv, verbose=True)
