"""seq_recommendations_b200: B200-native (sm_100a) next-item training / scoring hot path of
efikarra/seq-recommendations behind the reference's own model.py / preprocessor.py surface.

Importing the package is cheap and works without a GPU; constructing a model loads libseqrec_b200.so and a CUDA
device and raises if either is missing (there is no CPU fallback).
"""
from . import callbacks, optimizers, preprocessor  # noqa: F401
from ._lib import LIB_PATH, SeqrecError  # noqa: F401

__all__ = ["callbacks", "optimizers", "preprocessor", "LIB_PATH", "SeqrecError", "model", "engine", "synthetic"]


def __getattr__(name):
    if name in ("model", "engine", "synthetic", "dist"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
