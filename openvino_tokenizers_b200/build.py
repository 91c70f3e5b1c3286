"""Build libb200tok.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libb200tok.so"
SOURCES = ["api.cu", "tables.cpp", "regex_compile.cpp"]


def deps() -> list[Path]:
    """Every file the library is compiled from: all sources / headers / generated tables under csrc/ plus the public header."""
    out = [p for pat in ("*.cu", "*.cpp", "*.cuh", "*.hpp", "*.h", "*.inc") for p in CSRC.glob(pat)]
    out.append(CSRC.parent.parent / "include" / "b200tok.h")
    return out


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> Path:
    newest = max(p.stat().st_mtime for p in deps())
    if not force and LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v" if verbose else "-O3",
           "-cudart", "shared", "-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout)
        print(res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed")
    return LIB


if __name__ == "__main__":
    import sys
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
