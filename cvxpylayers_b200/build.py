"""Builds libbcone.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

Run as ``python -m cvxpylayers_b200.build``; ``__graft_entry__.build()`` calls :func:`build`.
nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SOURCES = ["api.cu", "fwd.cu", "fwd_fast.cu", "bwd.cu", "bwd_fast.cu", "bwd_block.cu", "pack.cu"]
HEADERS = [CSRC / "common.cuh", PKG.parent / "include" / "bcone.h"]
LIB = PKG / "libbcone.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; libbcone.so cannot be built")
    return cand


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [*(CSRC / s for s in SOURCES), *HEADERS])


def build(force: bool = False, verbose: bool = False, defines: tuple = (), out: "Path | None" = None) -> Path:
    """``defines`` / ``out``: development variants (e.g. ``-DBC_SUBPROF`` into another file, loaded with BCONE_LIB=...)."""
    target = Path(out) if out else LIB
    if not force and not defines and not out and not stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *defines, "-o", str(target), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return target


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=tuple(a for a in sys.argv[1:] if a.startswith("-D")), out=_out))
