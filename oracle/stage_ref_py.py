#!/usr/bin/env python
"""Stages the reference's PYTHON network sources for the drop-in test on the GPU box -> oracle/_ref/reference_py/

TEST INFRASTRUCTURE ONLY, like oracle/build_ref.py.  tests/test_dropin_network.py runs the reference's unmodified
networks (networks/DAIN.py, PWCNet/PWCNet.py, ...) on this package's operators and on the reference's own Python
layers (my_package/*/*Layer.py, *Module.py) over the reference's own kernels (oracle/_ref/*.so).  /root/reference does not exist on the GPU box, so the .py files the networks import are staged,
byte for byte, into the git-ignored oracle/_ref/ directory (never committed, never imported by the product, not
read by bench.py or smoke()); the test looks for /root/reference first and for this staged tree second.

    python oracle/stage_ref_py.py            (run where /root/reference exists; build() calls it)
"""
from __future__ import annotations

import os
import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("VFIDKR_REFERENCE", "/root/reference"))
OUT = HERE / "_ref" / "reference_py"

# what `import networks` pulls in (networks/DAIN.py:1-21, networks/DAIN_slowmotion.py:1-13)
TREES = ["networks", "S2D_models", "Resblock", "MegaDepth", "my_package"]   # my_package: its Python layers only
FILES = ["Stack.py", "PWCNet/__init__.py", "PWCNet/PWCNet.py",
         # the reference's own correlation wrapper, imported by PWCNet/PWCNet.py:15 when the aliases are NOT installed (run A)
         "PWCNet/correlation_package_pytorch1_0/__init__.py", "PWCNet/correlation_package_pytorch1_0/correlation.py"]


def available() -> bool:
    return (REF / "networks" / "DAIN.py").is_file()


def stage() -> Path:
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    if OUT.exists():
        shutil.rmtree(OUT)
    n = 0
    for tree in TREES:
        for src in sorted((REF / tree).rglob("*.py")):
            dst = OUT / src.relative_to(REF)
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(src, dst)
            n += 1
    for rel in FILES:
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / rel, dst)
        n += 1
    (OUT / "STAGED_FROM").write_text(f"{REF}\n{n} files, unmodified; test infrastructure, git-ignored\n")
    return OUT


def tree() -> Path | None:
    """The reference's Python tree to put on sys.path: the real one here, the staged copy on the GPU box."""
    if available():
        return REF
    if (OUT / "networks" / "DAIN.py").is_file():
        return OUT
    return None


if __name__ == "__main__":
    print(stage())
