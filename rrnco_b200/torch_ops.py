"""`torch.ops.rrnco_b200.*`: the C-ABI entry points registered with the PyTorch dispatcher (csrc/torch_ops.cpp ->
librrnco_b200_torch.so, built in-tree by rrnco_b200.build.build_torch_ops).  This is the default binding of the hot
calls when the library is present; `RRNCO_BINDING=ctypes` (or `use_torch_ops(False)`) selects the ctypes fallback."""
from __future__ import annotations

import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
TORCH_LIB_PATH = os.path.join(_HERE, "librrnco_b200_torch.so")
_state = {"loaded": False, "enabled": None}


def load() -> bool:
    """Load the op library once; False if it has not been built."""
    if _state["loaded"]:
        return True
    if not os.path.exists(TORCH_LIB_PATH):
        return False
    torch.ops.load_library(TORCH_LIB_PATH)
    _state["loaded"] = True
    return True


def enabled() -> bool:
    if _state["enabled"] is None:
        _state["enabled"] = os.environ.get("RRNCO_BINDING", "torch") != "ctypes" and load()
    return _state["enabled"]


def use_torch_ops(flag: bool = True) -> bool:
    """Switch the binding of the hot calls at run time; returns the binding now in effect (True = torch ops)."""
    _state["enabled"] = bool(flag) and load()
    return _state["enabled"]


def ops():
    if not load():
        raise RuntimeError(f"{TORCH_LIB_PATH} is missing: build it with `python -m rrnco_b200.build`")
    return torch.ops.rrnco_b200
