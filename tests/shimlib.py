"""Builds this repo's ov::Op shim against the stand-in OpenVINO API (tests/ov_stub) into
openvino_tokenizers_b200/csrc/ov_shim/libb200tok_ov_stub.so and drives it like oracle/ref.py drives the reference ops.
TEST INFRASTRUCTURE: the production build of the shim uses the real OpenVINO headers."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

from oracle import ref

ROOT = Path(__file__).resolve().parent.parent
SHIM_DIR = ROOT / "openvino_tokenizers_b200" / "csrc" / "ov_shim"
LIB = SHIM_DIR / "libb200tok_ov_stub.so"
SRC = [SHIM_DIR / "ov_extension_b200.cpp", ROOT / "tests" / "ov_stub" / "shim_driver.cpp"]
DEPS = SRC + [ROOT / "tests/ov_stub/stub_driver.hpp", ROOT / "tests/ov_stub/openvino/stub_core.hpp", ROOT / "include/b200tok.h"]


def build(force: bool = False) -> Path:
    from openvino_tokenizers_b200 import build as B
    core = B.build()
    newest = max(p.stat().st_mtime for p in DEPS)
    if force or not LIB.exists() or LIB.stat().st_mtime < max(newest, core.stat().st_mtime):
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-DIMPLEMENT_OPENVINO_EXTENSION_API", f"-I{ROOT / 'tests/ov_stub'}", f"-I{ROOT / 'include'}",
               "-o", str(LIB)] + [str(s) for s in SRC] + [f"-L{core.parent}", "-lb200tok", "-Wl,-rpath,$ORIGIN/..", "-Wl,--exclude-libs,ALL"]
        subprocess.check_call(cmd)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ref.lib(build(), prefix="ovshim")
    return _lib


class ShimOp(ref.StubOp):
    """One op of this repo's ov::Op shim, created the way the IR frontend creates a layer."""
    _prefix = "ovshim"

    def _lib(self):
        return lib()


class ShimGraph(ref.StubGraph):
    """An IR-like chain of layers loaded through this repo's extension entry point (incl. its load-time fusion)."""
    _prefix = "ovshim"

    def _lib(self):
        return lib()
