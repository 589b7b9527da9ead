"""Import the UNMODIFIED reference Python (as bytecode) on top of the UNMODIFIED reference kernels.

TEST INFRASTRUCTURE ONLY: used by tests/ and by `bench.py --impl reference`; never imported by `cnc_b200/`.

`oracle/build_ref.py:build_py()` byte-compiles the reference's hot-path modules from where they lie under
/root/reference into `oracle/_ref/py/*.pyc.bin` (binary build outputs, git-ignored, shipped to the GPU box next to the
reference `.so` files).  `load()` installs an import hook that resolves exactly the module names the reference's own
`import` statements use:

    utils, utils_bpp_acc, radiance_fields.ngp, datasets.utils, nerfacc(.grid/.scan/.volrend/.estimators.occ_grid ...)

and binds what those modules import from outside the reference tree:

    _gridencoder, pack_and_align, nerfacc.csrc  -> oracle/_ref/<name>/<name>.so  (reference CUDA, compiled unmodified)
    torchac        (third party, torchac==0.9.3, absent)   -> `_Torchac`: the oracle C range coder behind torchac's
                   encode_float_cdf / decode_float_cdf signatures ("parity unpinned")
    tinycudann     (third party, absent)                   -> `_Tcnn`: SphericalHarmonics degree 4, fp16 output
                   ("parity unpinned")

The reference modules create CUDA tensors at import (examples/utils.py:66-75, utils_bpp_acc.py:19-20), so `load()`
needs a GPU; `available()` says whether the build outputs are present at all.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

from . import ref_ext

_HERE = os.path.dirname(os.path.abspath(__file__))
PY_DIR = os.path.join(_HERE, "_ref", "py")
_PACKAGES = {"radiance_fields", "datasets", "nerfacc", "nerfacc.cuda", "nerfacc.estimators"}
_installed = False


def available() -> bool:
    need = ("utils", "utils_bpp_acc", "radiance_fields.ngp", "nerfacc", "nerfacc.estimators.occ_grid")
    return all(os.path.exists(os.path.join(PY_DIR, m + ".pyc.bin")) for m in need) and all(
        os.path.exists(ref_ext.path(n)) for n in ("_gridencoder", "pack_and_align", "nerfacc_csrc"))


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        pyc = os.path.join(PY_DIR, name + ".pyc.bin")
        is_pkg = name in _PACKAGES
        if os.path.exists(pyc):
            loader = importlib.machinery.SourcelessFileLoader(name, pyc)
            return importlib.util.spec_from_file_location(name, pyc, loader=loader,
                                                          submodule_search_locations=[PY_DIR] if is_pkg else None)
        if is_pkg:   # `datasets`: an empty package stands in for the image loaders
            return importlib.machinery.ModuleSpec(name, _Empty(), is_package=True)
        return None


class _Empty(importlib.abc.Loader):
    def create_module(self, spec):
        return None

    def exec_module(self, module):
        module.__path__ = [PY_DIR]


# ------------------------------------------------------------------------------------------ third-party stand-ins
def _make_torchac():
    """torchac 0.9.3's two entry points the reference calls (utils_bpp_acc.py:87,108) on the oracle coder.
    cdf_float [N, 3] fp32 on the CPU = [0, 1-p, 1]; int16 CDF = round(cdf * 65534) + arange(3)."""
    import numpy as np
    import torch

    from . import oracle as o

    m = types.ModuleType("torchac")

    def _c1(cdf_float):
        assert cdf_float.dim() == 2 and cdf_float.shape[1] == 3 and not cdf_float.is_cuda
        c = torch.round(cdf_float[:, 1].to(torch.float32) * 65534.0).to(torch.int32) + 1
        return (c.numpy().astype(np.int64) & 0xFFFF).astype(np.uint16)

    def encode_float_cdf(cdf_float, sym, needs_normalization=True, check_input_bounds=False):
        if check_input_bounds:
            assert float(cdf_float.min()) >= 0 and float(cdf_float.max()) <= 1
            assert int(sym.min()) >= 0 and int(sym.max()) <= 1
        return o.ac_encode(_c1(cdf_float), sym.numpy().astype(np.uint8))

    def decode_float_cdf(cdf_float, byte_stream, needs_normalization=True):
        return torch.from_numpy(o.ac_decode(_c1(cdf_float), byte_stream).astype(np.int16))

    m.encode_float_cdf, m.decode_float_cdf = encode_float_cdf, decode_float_cdf
    return m


def _make_tcnn():
    import torch

    m = types.ModuleType("tinycudann")

    class Encoding(torch.nn.Module):
        """Composite[SphericalHarmonics degree 4] (ngp.py:412-425): input in [0,1]^3, 16 fp16 outputs"""

        def __init__(self, n_input_dims, encoding_config, **kw):
            super().__init__()
            nested = encoding_config["nested"]
            assert len(nested) == 1 and nested[0]["otype"] == "SphericalHarmonics" and nested[0]["degree"] == 4
            self.n_input_dims, self.n_output_dims = n_input_dims, 16

        def forward(self, d01):
            x, y, z = (d01.float() * 2 - 1).unbind(-1)
            xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
            o = [torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z,
                 -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
                 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
                 0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
                 2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
                 0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
                 1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)]
            return torch.stack(o, -1).half()

    m.Encoding = Encoding
    return m


def load():
    """install the hook and the bindings; returns a namespace with the reference modules:
    .ngp (radiance_fields.ngp), .bpp (utils_bpp_acc), .utils (examples/utils.py), .nerfacc"""
    global _installed
    if not available():
        raise RuntimeError("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("the reference modules create CUDA tensors at import: a GPU is required")
    if not _installed:
        # the reference's top-level names are generic (`utils`, `datasets`): whatever else answers to them must go
        for name in [m for m in sys.modules if m.split(".")[0] in ("utils", "utils_bpp_acc", "datasets", "nerfacc", "radiance_fields",
                                                                   "torchac", "tinycudann", "_gridencoder", "pack_and_align")]:
            del sys.modules[name]
        for name in ("_gridencoder", "pack_and_align"):
            sys.modules[name] = ref_ext.load(name)
        sys.modules["torchac"] = _make_torchac()
        sys.modules["tinycudann"] = _make_tcnn()
        sys.meta_path.insert(0, _Finder())
        _installed = True
        import nerfacc   # noqa: F401  (resolved by _Finder: the reference's vendored copy)

        sys.modules["nerfacc.csrc"] = ref_ext.load("nerfacc_csrc")
        nerfacc.csrc = sys.modules["nerfacc.csrc"]
    import nerfacc
    import utils
    import utils_bpp_acc
    from radiance_fields import ngp

    import datasets.utils as du

    for m in (nerfacc, utils, utils_bpp_acc, ngp, du):
        assert m.__spec__.origin.startswith(PY_DIR), f"another `{m.__name__}` shadows the reference's"
    return types.SimpleNamespace(ngp=ngp, bpp=utils_bpp_acc, utils=utils, nerfacc=nerfacc)
