"""ctypes binding of include/idn_host.h (libidn_host.so): the C++ host mirror of idencomp's API.

Model loading (msgpack + SHA3 identifier check + bit-exact f32 quantiser) happens in the C++ library; Python only
passes paths and pointers.  No CPU codec lives here: everything that touches symbols goes to libidn_gpu.so.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libidn_host.so"

EXPORTS = [
    "idn_host_last_error", "idn_host_model_load", "idn_host_model_from_bytes", "idn_host_model_new",
    "idn_host_model_empty", "idn_host_model_free", "idn_host_model_type", "idn_host_model_len",
    "idn_host_model_spec_name", "idn_host_model_identifier", "idn_host_model_cum_table", "idn_host_model_upload",
    "idn_host_quantise",
]

_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    capi.load()
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(str(LIB_PATH))
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
    L.idn_host_last_error.restype = C.c_char_p
    L.idn_host_model_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.idn_host_model_from_bytes.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.idn_host_model_new.argtypes = [i32, C.c_char_p, u32, vp, vp, vp, u64, C.POINTER(vp)]
    L.idn_host_model_empty.argtypes = [i32, C.POINTER(vp)]
    L.idn_host_model_free.argtypes = [vp]
    L.idn_host_model_free.restype = None
    L.idn_host_model_type.argtypes = [vp]
    L.idn_host_model_len.argtypes = [vp]
    L.idn_host_model_len.restype = u32
    L.idn_host_model_spec_name.argtypes = [vp]
    L.idn_host_model_spec_name.restype = C.c_char_p
    L.idn_host_model_identifier.argtypes = [vp, vp]
    L.idn_host_model_identifier.restype = None
    L.idn_host_model_cum_table.argtypes = [vp, vp, u64]
    L.idn_host_model_cum_table.restype = u64
    L.idn_host_model_upload.argtypes = [vp, vp, C.POINTER(i32)]
    L.idn_host_quantise.argtypes = [vp, u32, u32, vp]
    _LIB = L
    return L


class HostError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{capi.ERRORS.get(code, code)}: {msg}")
        self.code = code
        self.kind = capi.ERRORS.get(code, str(code))


def _check(rc: int):
    if rc != 0:
        raise HostError(rc, load().idn_host_last_error().decode())


def quantise(probs, scale_bits: int = 14) -> np.ndarray:
    """Context::as_integer_cum_freqs (context.rs:346-371)."""
    p = np.ascontiguousarray(probs, dtype=np.float32)
    out = np.zeros(len(p), dtype=np.uint32)
    _check(load().idn_host_quantise(p.ctypes.data, len(p), scale_bits, out.ctypes.data))
    return out


class Model:
    """idencomp::Model (model.rs:176-259)."""

    def __init__(self, handle):
        self.L = load()
        self.h = handle

    @staticmethod
    def load(path) -> "Model":
        h = C.c_void_p()
        _check(load().idn_host_model_load(str(path).encode(), C.byref(h)))
        return Model(h)

    @staticmethod
    def from_bytes(data: bytes) -> "Model":
        h = C.c_void_p()
        buf = np.frombuffer(data, dtype=np.uint8)
        _check(load().idn_host_model_from_bytes(buf.ctypes.data, len(buf), C.byref(h)))
        return Model(h)

    @staticmethod
    def new(model_type: int, spec_name: str, probs, spec_keys, spec_ctx) -> "Model":
        p = np.ascontiguousarray(probs, dtype=np.float32)
        k = np.ascontiguousarray(spec_keys, dtype=np.uint32)
        c = np.ascontiguousarray(spec_ctx, dtype=np.uint32)
        h = C.c_void_p()
        _check(load().idn_host_model_new(model_type, spec_name.encode(), p.shape[0] if p.ndim == 2 else 0,
                                         p.ctypes.data if p.size else None, k.ctypes.data if k.size else None,
                                         c.ctypes.data if c.size else None, len(k), C.byref(h)))
        return Model(h)

    @staticmethod
    def empty(model_type: int) -> "Model":
        h = C.c_void_p()
        _check(load().idn_host_model_empty(model_type, C.byref(h)))
        return Model(h)

    def __del__(self):
        try:
            if self.h:
                self.L.idn_host_model_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def model_type(self) -> int:
        return int(self.L.idn_host_model_type(self.h))

    def __len__(self) -> int:
        return int(self.L.idn_host_model_len(self.h))

    @property
    def spec_name(self) -> str:
        return self.L.idn_host_model_spec_name(self.h).decode()

    @property
    def identifier(self) -> bytes:
        out = (C.c_uint8 * 32)()
        self.L.idn_host_model_identifier(self.h, out)
        return bytes(out)

    def cum_table(self) -> np.ndarray:
        nsym = 5 if self.model_type == 0 else 94
        out = np.zeros((len(self) + 1, nsym + 1), dtype=np.uint16)
        self.L.idn_host_model_cum_table(self.h, out.ctypes.data, out.size)
        return out

    def upload(self, ctx: "capi.Context") -> int:
        """RansEncModel/RansDecModel::from_model on the device; returns the idn_model_t handle."""
        h = C.c_int32(-1)
        _check(self.L.idn_host_model_upload(ctx.h, self.h, C.byref(h)))
        return int(h.value)
