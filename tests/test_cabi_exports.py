"""CPU suite: the C-ABI library builds, loads and exports every symbol include/cnc_b200.h
declares; the Python shims refuse CPU tensors loudly (no fallback).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "cnc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cnc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from cnc_b200 import build, _lib

    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/cnc_b200.h but not exported"
    # every compute entry point has a ctypes signature in the shim
    for n in names:
        if n not in ("cnc_version", "cnc_last_error", "cnc_field_blob_floats", "cnc_context3d_mlp_floats",
                     "cnc_wgrad_max_partials", "cnc_dgrad_blob_floats", "cnc_ctx_mlp_floats", "cnc_ctx_mlp_max_partials",
                         "cnc_lin8_rows_per_block", "cnc_peer_handle_bytes", "cnc_peer_pad_bytes",
                         "cnc_bernoulli_bits_blocks"):
            assert n in _lib.SIGNATURES, n
    L.cnc_version.restype = ctypes.c_int
    assert L.cnc_version() >= 100


def test_shims_fail_loudly_without_cuda_tensors():
    from cnc_b200 import _gridencoder, pack_and_align

    x = torch.rand(4, 3)
    t = torch.rand(64, 2)
    o = torch.tensor([0, 64], dtype=torch.int32)
    r = torch.tensor([4], dtype=torch.int32)
    out = torch.empty(1, 4, 2)
    with pytest.raises(RuntimeError):
        _gridencoder.grid_encode_forward(x, t, o, r, out, 4, 3, 2, 1, 0, 128, 0)
    with pytest.raises(RuntimeError):
        pack_and_align.query_mask_3D(torch.zeros(4, 3, dtype=torch.int16), torch.zeros(8, 8, 8, dtype=torch.bool),
                                     torch.zeros(4, dtype=torch.int16), torch.zeros(4, dtype=torch.int32), 18, 4)


def test_missing_library_is_an_error(monkeypatch):
    from cnc_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcnc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cnc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libcnc_oracle" not in src, f


def test_grid_encoder_layout_matches_reference(golden):
    from cnc_b200.gridencoder import GridEncoder
    from conftest import R2, R3

    e = GridEncoder(num_dim=3, n_features=8, resolutions_list=R3, log2_hashmap_size=19, ste_binary=True)
    assert e.offsets_list.dtype == torch.int32 and e.resolutions_list.dtype == torch.int32
    assert e.offsets_list.tolist() == golden["layout_xyz_offsets"].tolist()
    assert tuple(e.params.shape) == (4003896, 8) and e.n_output_dims == 96
    assert float(e.params.abs().max()) <= 1e-4
    p = GridEncoder(num_dim=2, n_features=8, resolutions_list=R2, log2_hashmap_size=17, ste_binary=True)
    assert p.offsets_list.tolist() == golden["layout_plane_offsets"].tolist()
