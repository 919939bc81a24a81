"""ctypes loader for oracle/c/oracle.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        try:
            build()
            _lib = ctypes.CDLL(_SO)
        except Exception:  # no compiler: callers fall back to the numpy loops
            _lib = False
    return _lib


def available() -> bool:
    return bool(_load())


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def gae_return(value, value_next, rew, end_flag, gamma, gae_lambda):
    lib = _load()
    value = np.ascontiguousarray(value, dtype=np.float32)
    value_next = np.ascontiguousarray(value_next, dtype=np.float32)
    rew = np.ascontiguousarray(rew, dtype=np.float64)
    end = np.ascontiguousarray(np.asarray(end_flag) != 0, dtype=np.uint8)
    n = rew.shape[0]
    out = np.zeros(n, dtype=np.float64)
    lib.oracle_gae_return(_p(value, ctypes.c_float), _p(value_next, ctypes.c_float),
                          _p(rew, ctypes.c_double), _p(end, ctypes.c_uint8),
                          ctypes.c_double(gamma), ctypes.c_double(gae_lambda),
                          ctypes.c_int64(n), _p(out, ctypes.c_double))
    return out


def nstep_return(metric, end_flag, target_q, indices, gamma, n_step):
    lib = _load()
    metric = np.ascontiguousarray(metric, dtype=np.float64)
    end = np.ascontiguousarray(np.asarray(end_flag) != 0, dtype=np.uint8)
    shape = np.asarray(target_q).shape
    bsz = shape[0]
    tq = np.ascontiguousarray(np.asarray(target_q, dtype=np.float64).reshape(bsz, -1)).copy()
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    lib.oracle_nstep_return(_p(metric, ctypes.c_double), _p(end, ctypes.c_uint8),
                            _p(tq, ctypes.c_double), _p(idx, ctypes.c_int64),
                            ctypes.c_double(gamma), ctypes.c_int64(n_step),
                            ctypes.c_int64(bsz), ctypes.c_int64(tq.shape[1]))
    return tq.reshape(shape)
