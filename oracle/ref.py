"""ctypes driver for oracle/_ref/libovtok_ref.so — the REFERENCE's own op classes (RegexSplit, BPETokenizer,
WordpieceTokenizer, VocabEncoder, VocabDecoder, ByteFallback, SpecialTokensSplit, Truncate, CombineSegments, RaggedToDense,
BytesToChars, CharsToBytes, FuzeRagged, UTF8Validate, RegexNormalization) compiled unmodified from /root/reference/src
against the stand-in OpenVINO API (tests/ov_stub).  TEST INFRASTRUCTURE ONLY (same rule as oracle.py): only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import it.

``RefOp(name, protos, **attrs)`` "loads a layer" the way the IR frontend does (default-construct, connect inputs,
visit_attributes, validate_and_infer_types); ``op(*arrays)`` calls the reference's ``evaluate()`` on host tensors that
view the numpy memory and returns copies of the outputs.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libovtok_ref.so"
REFERENCE_SRC = Path("/root/reference/src")

_CODES = {np.dtype(np.int32): 1, np.dtype(np.int64): 2, np.dtype(np.uint8): 3, np.dtype(np.bool_): 4, np.dtype(np.float32): 5}
_DTYPES = {v: k for k, v in _CODES.items()}


class _Tensor(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int64 * 4), ("data", C.c_void_p)]


def build(force: bool = False) -> Path | None:
    """Compile oracle/_ref with the committed Makefile where the reference sources exist (this container); elsewhere
    (the GPU box) the prebuilt library travels with the snapshot.  Returns None when neither is available."""
    if REFERENCE_SRC.exists():
        deps = [REFERENCE_SRC / "bpe_tokenizer.cpp", _HERE / "ref_driver.cpp", _HERE.parent / "tests/ov_stub/stub_driver.hpp",
                _HERE.parent / "tests/ov_stub/openvino/stub_core.hpp", _HERE / "Makefile"]
        if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(d.stat().st_mtime for d in deps):
            subprocess.check_call(["make", "-C", str(_HERE), "-s", "-j8", "_ref"])
    return LIB_PATH if LIB_PATH.exists() else None


def available() -> bool:
    return build() is not None


_lib = None


def lib(path: Path | None = None, prefix: str = "ovref"):
    global _lib
    if path is None and _lib is not None:
        return _lib
    p = path or build()
    if p is None:
        raise RuntimeError("oracle/_ref/libovtok_ref.so is missing and /root/reference is not present to build it")
    L = C.CDLL(str(p))
    getattr(L, f"{prefix}_node_create").restype = C.c_void_p
    getattr(L, f"{prefix}_node_create").argtypes = [C.c_char_p, C.c_int, C.POINTER(_Tensor), C.c_char_p]
    getattr(L, f"{prefix}_node_evaluate").argtypes = [C.c_void_p, C.c_int, C.POINTER(_Tensor)]
    getattr(L, f"{prefix}_node_n_outputs").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_node_output").argtypes = [C.c_void_p, C.c_int, C.POINTER(_Tensor)]
    getattr(L, f"{prefix}_node_last_ms").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_node_last_ms").restype = C.c_double
    getattr(L, f"{prefix}_node_type").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_node_type").restype = C.c_char_p
    getattr(L, f"{prefix}_node_share").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_node_share").restype = C.c_void_p
    getattr(L, f"{prefix}_node_destroy").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_last_error").restype = C.c_char_p
    getattr(L, f"{prefix}_graph_create").restype = C.c_void_p
    getattr(L, f"{prefix}_graph_destroy").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_graph_add_input").argtypes = [C.c_void_p, C.POINTER(_Tensor)]
    getattr(L, f"{prefix}_graph_add_layer").argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_int), C.c_int]
    getattr(L, f"{prefix}_graph_value_type").argtypes = [C.c_void_p, C.c_int]
    getattr(L, f"{prefix}_graph_value_type").restype = C.c_char_p
    getattr(L, f"{prefix}_graph_evaluate").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_Tensor)]
    getattr(L, f"{prefix}_graph_evaluated_nodes").argtypes = [C.c_void_p]
    getattr(L, f"{prefix}_graph_result").argtypes = [C.c_void_p, C.c_int, C.POINTER(_Tensor)]
    if path is None:
        _lib = L
    return L


def _desc(a: np.ndarray, with_data: bool) -> _Tensor:
    t = _Tensor()
    t.dtype = _CODES[a.dtype]
    t.ndim = a.ndim
    for i, s in enumerate(a.shape):
        t.shape[i] = s
    t.data = a.ctypes.data if (with_data and a.size) else None
    return t


def as_tensor(a) -> np.ndarray:
    """Arrays exactly as the ops see them: bytes / str become u8[len]; Python ints become i32 scalars."""
    if isinstance(a, str):
        a = a.encode()
    if isinstance(a, (bytes, bytearray)):
        return np.frombuffer(bytes(a), dtype=np.uint8).copy()
    if isinstance(a, (int, np.integer)) and not isinstance(a, np.generic):
        return np.asarray(a, np.int32)
    a = np.asarray(a)
    if a.dtype not in _CODES:
        raise TypeError(f"unsupported dtype {a.dtype}")
    return a if a.ndim == 0 else np.ascontiguousarray(a)      # (ascontiguousarray would turn a scalar into shape [1])


class StubOp:
    """One op instance driven through the stand-in runtime (base for the reference ops and for this repo's shim)."""

    _prefix = "ovref"

    def _lib(self):
        return lib()

    def __init__(self, name: str, n_inputs_or_protos, constants: dict[int, np.ndarray] | None = None, **attrs):
        """``n_inputs_or_protos``: list of example arrays (dtype + rank are used); ``constants``: {input index: array}
        inputs that are Constant nodes in the IR (vocab, merges, patterns) — the others are Parameters."""
        L = self._lib()
        protos = [as_tensor(a) for a in n_inputs_or_protos]
        constants = {k: as_tensor(v) for k, v in (constants or {}).items()}
        arr = (_Tensor * len(protos))()
        self._keep = []
        for i, a in enumerate(protos):
            if i in constants:
                arr[i] = _desc(constants[i], True)
                if constants[i].size == 0:      # an empty Constant still has to be a Constant: point at a dummy byte
                    dummy = np.zeros(1, np.uint8)
                    self._keep.append(dummy)
                    arr[i].data = dummy.ctypes.data
            else:
                arr[i] = _desc(a, False)
        text = "\x1e".join(f"{k}={self._fmt(v)}" for k, v in attrs.items()).encode("utf-8", "surrogateescape")
        self._h = getattr(L, f"{self._prefix}_node_create")(name.encode(), len(protos), arr, text)
        if not self._h:
            raise RuntimeError(f"{name}: " + getattr(L, f"{self._prefix}_last_error")().decode(errors="replace"))
        self.name = name
        self.node_type = getattr(L, f"{self._prefix}_node_type")(self._h).decode()
        self.last_ms = 0.0

    def share(self):
        """Another handle on the same op instance (shared tables / caches), for concurrent evaluate() calls from several threads —
        what OpenVINO does with several infer requests on one compiled model."""
        other = object.__new__(type(self))
        other._keep, other.name, other.node_type, other.last_ms = self._keep, self.name, self.node_type, 0.0
        other._h = getattr(self._lib(), f"{self._prefix}_node_share")(self._h)
        return other

    @staticmethod
    def _fmt(v):
        if isinstance(v, bool):
            return "true" if v else "false"
        if isinstance(v, bytes):
            return v.decode("utf-8", "surrogateescape")
        if isinstance(v, (list, tuple, np.ndarray)):
            return ", ".join(str(int(x)) for x in v)
        return str(v)

    def __call__(self, *inputs):
        L = self._lib()
        arrs = [as_tensor(a) for a in inputs]
        tin = (_Tensor * len(arrs))()
        for i, a in enumerate(arrs):
            tin[i] = _desc(a, True)
            if a.ndim == 0:             # scalars always carry data
                tin[i].data = a.ctypes.data
        rc = getattr(L, f"{self._prefix}_node_evaluate")(self._h, len(arrs), tin)
        if rc:
            raise RuntimeError(f"{self.name}.evaluate: " + getattr(L, f"{self._prefix}_last_error")().decode(errors="replace"))
        self.last_ms = getattr(L, f"{self._prefix}_node_last_ms")(self._h)
        outs = []
        for i in range(getattr(L, f"{self._prefix}_node_n_outputs")(self._h)):
            d = _Tensor()
            getattr(L, f"{self._prefix}_node_output")(self._h, i, C.byref(d))
            shape = tuple(d.shape[k] for k in range(d.ndim))
            n = int(np.prod(shape)) if shape else 1
            dt = _DTYPES.get(d.dtype)
            if dt is None:
                outs.append(None)
                continue
            if n == 0 or not d.data:
                outs.append(np.zeros(shape, dt))
                continue
            buf = (C.c_uint8 * (n * dt.itemsize)).from_address(d.data)
            outs.append(np.frombuffer(buf, dtype=dt).reshape(shape).copy())
        return outs

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                getattr(self._lib(), f"{self._prefix}_node_destroy")(h)
            except Exception:
                pass


class RefOp(StubOp):
    """A reference op class (src/*.cpp of openvino_tokenizers) — evaluate() is the reference's own code."""


def _tensor_to_numpy(d: _Tensor):
    shape = tuple(d.shape[k] for k in range(d.ndim))
    n = int(np.prod(shape)) if shape else 1
    dt = _DTYPES.get(d.dtype)
    if dt is None:
        return None
    if n == 0 or not d.data:
        return np.zeros(shape, dt)
    buf = (C.c_uint8 * (n * dt.itemsize)).from_address(d.data)
    return np.frombuffer(buf, dtype=dt).reshape(shape).copy()


class StubGraph:
    """A chain of layers "read" one after the other through the registered extensions — what loading an IR does — with a minimal
    executor.  ``parameter(example)`` / ``constant(array)`` return value ids; ``layer(op, inputs, **attrs)`` returns the ids of the
    layer's outputs; ``run(wanted_ids, parameter_arrays)`` evaluates the producing nodes in dependency order."""

    _prefix = "ovref"

    def _lib(self):
        return lib()

    def _fn(self, name):
        return getattr(self._lib(), f"{self._prefix}_{name}")

    def __init__(self):
        self._g = self._fn("graph_create")()
        self._keep = []

    def _err(self, what):
        return RuntimeError(f"{what}: " + self._fn("last_error")().decode(errors="replace"))

    def parameter(self, example):
        t = _desc(as_tensor(example), False)
        i = self._fn("graph_add_input")(self._g, C.byref(t))
        if i < 0:
            raise self._err("parameter")
        return i

    def constant(self, value):
        a = as_tensor(value)
        self._keep.append(a)
        t = _desc(a, True)
        if a.size == 0:
            dummy = np.zeros(1, np.uint8)
            self._keep.append(dummy)
            t.data = dummy.ctypes.data
        if a.ndim == 0:
            t.data = a.ctypes.data
        i = self._fn("graph_add_input")(self._g, C.byref(t))
        if i < 0:
            raise self._err("constant")
        return i

    def layer(self, op, inputs, **attrs):
        ids = (C.c_int * len(inputs))(*inputs)
        out = (C.c_int * 16)()
        text = "\x1e".join(f"{k}={StubOp._fmt(v)}" for k, v in attrs.items()).encode("utf-8", "surrogateescape")
        n = self._fn("graph_add_layer")(self._g, op.encode(), len(inputs), ids, text, out, 16)
        if n < 0:
            raise self._err(op)
        return [out[i] for i in range(n)]

    def producer_type(self, value_id):
        return self._fn("graph_value_type")(self._g, value_id).decode()

    def run(self, wanted, parameter_arrays):
        arrs = [as_tensor(a) for a in parameter_arrays]
        tin = (_Tensor * max(len(arrs), 1))()
        for i, a in enumerate(arrs):
            tin[i] = _desc(a, True)
            if a.ndim == 0:
                tin[i].data = a.ctypes.data
        want = (C.c_int * len(wanted))(*wanted)
        if self._fn("graph_evaluate")(self._g, len(wanted), want, len(arrs), tin):
            raise self._err("evaluate")
        res = []
        for i in range(len(wanted)):
            d = _Tensor()
            self._fn("graph_result")(self._g, i, C.byref(d))
            res.append(_tensor_to_numpy(d))
        return res

    @property
    def evaluated_nodes(self):
        return self._fn("graph_evaluated_nodes")(self._g)

    def __del__(self):
        g, self._g = getattr(self, "_g", None), None
        if g:
            try:
                self._fn("graph_destroy")(g)
            except Exception:
                pass
