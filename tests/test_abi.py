"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/fsrl_b200.h declares (no compute calls -- there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fsrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsrl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fsrl_b200 import _lib
    names = _declared()
    assert len(names) >= 5
    for n in names:
        assert hasattr(_lib.lib, n), f"libfsrl_b200.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in _lib.SIGNATURES"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in python but not declared in include/fsrl_b200.h"


def test_abi_version_and_error_text():
    from fsrl_b200 import _lib
    assert _lib.lib.fsrl_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_argument_validation_needs_no_gpu():
    """EINVAL paths return before touching the device."""
    from fsrl_b200 import _lib
    rc = _lib.lib.fsrl_gae_dual(None, None, None, None, None, None, 0.99, 0.95, None, None,
                                -1, 0, 2, None, 0, None)
    assert rc == _lib.FSRL_EINVAL and "N must be" in _lib.last_error()
    rc = _lib.lib.fsrl_gae_dual(None, None, None, None, None, None, 0.99, 1.5, None, None,
                                8, 8, 2, None, 0, None)
    assert rc == _lib.FSRL_EINVAL and "GAE lambda should be in [0, 1]." in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert _lib.lib.fsrl_gae_dual_workspace_bytes(614400) > 0


def test_ops_refuse_cpu_tensors():
    import torch
    from fsrl_b200 import ops
    x = torch.zeros(2, 8)
    with pytest.raises(TypeError, match="CUDA tensor"):
        ops.gae_dual(x, x, x[0], x[0], torch.zeros(8, dtype=torch.uint8), None, 0.99, 0.95)
