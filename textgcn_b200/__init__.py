"""textgcn_b200 — sm_100a kernels beneath TextGCN's LightGCN hot path, behind the reference's model API.

Importing the package does not load the CUDA library; the first compute call does, and raises if
``libtgcn_b200.so`` has not been built (there is no CPU fallback).
"""
from ._lib import LIB_PATH, TgcnError  # noqa: F401

__all__ = ["BaseModel", "AdvSamplModel", "LTRLinear", "LTRLinearWPop", "B200HotPath", "make_params", "TgcnError"]


def __getattr__(name):
    if name in __all__:
        from . import models
        return getattr(models, name)
    raise AttributeError(name)
