"""ctypes binding of libfsrl_b200.so (the C-ABI in include/fsrl_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a symbol is absent the
import fails loudly (a silent CPU path would void every parity claim).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsrl_b200.so")

FSRL_OK, FSRL_EINVAL, FSRL_ECUDA, FSRL_EWORKSPACE = 0, -1, -2, -3


class FsrlCudaError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C fsrl_b200/csrc`). fsrl_b200 has no CPU fallback.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

c_f32p = ctypes.c_void_p   # device pointers travel as integers
c_u8p = ctypes.c_void_p
c_i32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
c_f32 = ctypes.c_float
c_int = ctypes.c_int
c_size = ctypes.c_size_t
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "fsrl_last_error": (ctypes.c_char_p, []),
    "fsrl_abi_version": (c_int, []),
    "fsrl_sm_count": (c_int, []),
    "fsrl_gae_dual_workspace_bytes": (c_size, [c_i64]),
    "fsrl_gae_dual": (c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_u8p, c_u8p, c_f64, c_f64,
                              c_f32p, c_f32p, c_i64, c_i64, c_int, c_vp, c_size, c_vp]),
}


def _bind():
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover - build error
            raise ImportError(f"libfsrl_b200.so lacks symbol {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args


_bind()


def last_error() -> str:
    return lib.fsrl_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Translate a C-ABI return code into the exception the reference would raise."""
    if rc == FSRL_OK:
        return
    msg = last_error()
    if rc == FSRL_EINVAL:
        raise ValueError(msg)
    if rc == FSRL_EWORKSPACE:
        raise MemoryError(msg)
    raise FsrlCudaError(msg)
