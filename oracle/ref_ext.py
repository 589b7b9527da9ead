"""Loader for the prebuilt reference extensions in oracle/_ref (TEST INFRASTRUCTURE ONLY).

`load("_gridencoder")`, `load("pack_and_align")`, `load("nerfacc_csrc")` return the pybind module
compiled by oracle/build_ref.py from the unmodified reference sources, or None when the `.so`
is absent (then the GPU A/B tests skip).  Never imported by the product package.
"""
from __future__ import annotations

import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def path(name: str) -> str:
    return os.path.join(_HERE, "_ref", name, f"{name}.so")


def load(name: str):
    if name in _cache:
        return _cache[name]
    p = path(name)
    mod = None
    if os.path.exists(p):
        import torch  # noqa: F401  (libtorch symbols must be loaded first)

        spec = importlib.util.spec_from_file_location(name, p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod
