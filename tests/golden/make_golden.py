"""Generate tests/golden/golden_v1.npz by *running the reference's own Python* on CPU.

Run in the build container only (needs /root/reference); the GPU box and the test-suite
read the committed .npz.  Nothing from the reference is copied into the repo: the
functions are pulled out of the reference files with `ast` at run time and executed
in a scratch namespace (the modules themselves cannot be imported here: they create
CUDA tensors / import tinycudann, torchac, _gridencoder at import time -- SURVEY F2, 8c).

Functions executed (reference file:line):
  get_grid_index                     examples/utils.py:492-511
  STE_binary, GridEncoder.__init__   examples/radiance_fields/ngp.py:22-39, 171-223
  Embedder / get_embedder, _TruncExp examples/radiance_fields/ngp.py:318-334, 569-617
  Bernoulli_entropy                  examples/utils_bpp_acc.py:1002-1013
Plus the nerfacc docstring known-answer vectors, typed from
  nerfacc/pack.py:29-32, nerfacc/scan.py:36-39,78-81,127-130,170-173,
  nerfacc/volrend.py:194-197,248-255,300-304,349-357,405-411,463-473.
"""
import ast
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CNC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")


def pull(path, names, ns):
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = compile(ast.Module([node], []), path, "exec")
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    import torch.nn as nn
    from torch.autograd import Function

    def _ident_deco(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    ns = dict(torch=torch, np=np, nn=nn, Function=Function, custom_fwd=_ident_deco,
              custom_bwd=_ident_deco)
    pull("examples/utils.py", ["get_grid_index"], ns)
    pull("examples/radiance_fields/ngp.py",
         ["STE_binary", "GridEncoder", "Embedder", "get_embedder", "_TruncExp"], ns)
    pull("examples/utils_bpp_acc.py", ["Bernoulli_entropy"], ns)
    g = {}
    rng = np.random.default_rng(1234)

    # ---- a1: hash / dense index, int32 and int64 callers -----------------------
    R3 = [18, 24, 33, 44, 59, 80, 108, 148, 201, 275, 376, 514]
    R2 = [130, 258, 514, 1026]
    R16 = [int(np.floor(16 * (512 / 16) ** (l / 15))) + 2 for l in range(16)]
    cases = []
    for D, res_list, log2T in ((3, R3, 19), (2, R2, 17), (3, R16, 14)):
        for res in res_list:
            T = min(2 ** log2T, res ** D)
            T = int(np.ceil(T / 8) * 8)
            cases.append((D, res, T))
    g["hash_cases"] = np.array(cases, np.int64)
    for k, (D, res, T) in enumerate(cases):
        pos = rng.integers(0, res, size=(257, D)).astype(np.int64)
        pos[0] = 0
        pos[1] = res - 1
        out32 = ns["get_grid_index"](T, res, torch.from_numpy(pos.astype(np.int32)).unsqueeze(1))[:, 0]
        out64 = ns["get_grid_index"](T, res, torch.from_numpy(pos).unsqueeze(1))[:, 0]
        assert torch.equal(out32, out64), "int32/int64 callers disagree (non power-of-two T?)"
        g[f"hash_pos_{k}"] = pos
        g[f"hash_idx_{k}"] = out64.numpy().astype(np.int64)

    # ---- table layout of GridEncoder.__init__ -----------------------------------------------
    for name, (D, res_list, log2T, F) in dict(
        xyz=(3, R3, 19, 8), plane=(2, R2, 17, 8), cfg1=(3, R16, 14, 2)
    ).items():
        enc = ns["GridEncoder"](num_dim=D, n_features=F, resolutions_list=res_list,
                                log2_hashmap_size=log2T, ste_binary=True)
        g[f"layout_{name}_offsets"] = enc.offsets_list.numpy().astype(np.int32)
        g[f"layout_{name}_res"] = enc.resolutions_list.numpy().astype(np.int32)
        g[f"layout_{name}_rows"] = np.array(enc.params.shape[0], np.int64)

    # ---- STE_binary fwd / bwd -----------------------------------------------------------------
    v = np.concatenate([rng.normal(0, 1.2, 200), [0.0, -0.0, 1.0, -1.0, 1.0000001, -1.0000001, 1e-30, -1e-30]])
    t = torch.tensor(v, dtype=torch.float32, requires_grad=True)
    y = ns["STE_binary"].apply(t)
    gr = torch.tensor(rng.normal(size=t.shape), dtype=torch.float32)
    y.backward(gr)
    g["ste_in"], g["ste_out"] = t.detach().numpy(), y.detach().numpy()
    g["ste_gin"], g["ste_gout"] = gr.numpy(), t.grad.numpy()

    # ---- frequency embedding + trunc_exp + Bernoulli entropy ------------------------------------
    embed, ch = ns["get_embedder"](10, 0)
    x = torch.tensor(rng.uniform(0, 1, (64, 3)), dtype=torch.float32)
    g["embed_in"], g["embed_out"] = x.numpy(), embed(x).numpy()
    assert ch == 63
    h = torch.tensor(rng.normal(0, 3, 64), dtype=torch.float32)
    g["texp_in"], g["texp_out"] = h.numpy(), ns["_TruncExp"].apply(h - 1).numpy()
    be = ns["Bernoulli_entropy"]()
    xs = torch.tensor(rng.integers(0, 2, 128) * 2.0 - 1.0, dtype=torch.float32)
    ps = torch.tensor(np.concatenate([rng.uniform(-0.2, 1.2, 120), [0, 1, 1e-7, 1 - 1e-7, 0.5, 0.25, 1e-6, 1 - 1e-6]]),
                      dtype=torch.float32)
    g["bern_x"], g["bern_p"], g["bern_bits"] = xs.numpy(), ps.numpy(), be(xs, ps).numpy()

    # ---- nerfacc docstring KATs -------------------------------------------------------------------
    g["kat_ray_indices_9"] = np.array([0, 0, 1, 1, 1, 2, 2, 2, 2], np.int64)
    g["kat_packed_info"] = np.array([[0, 2], [2, 3], [5, 4]], np.int64)
    g["kat_scan_in"] = np.arange(1, 10, dtype=np.float32)
    g["kat_inclusive_sum"] = np.array([1, 3, 3, 7, 12, 6, 13, 21, 30], np.float32)
    g["kat_exclusive_sum"] = np.array([0, 1, 0, 3, 7, 0, 6, 13, 21], np.float32)
    g["kat_inclusive_prod"] = np.array([1, 2, 3, 12, 60, 6, 42, 336, 3024], np.float32)
    g["kat_exclusive_prod"] = np.array([1, 1, 1, 3, 12, 1, 6, 42, 336], np.float32)
    g["kat_alphas"] = np.array([0.4, 0.8, 0.1, 0.8, 0.1, 0.0, 0.9], np.float32)
    g["kat_ray_indices_7"] = np.array([0, 0, 0, 1, 1, 2, 2], np.int64)
    g["kat_trans_from_alpha"] = np.array([1.0, 0.6, 0.12, 1.0, 0.2, 1.0, 1.0], np.float32)
    g["kat_weights_from_alpha"] = np.array([0.4, 0.48, 0.012, 0.8, 0.02, 0.0, 0.9], np.float32)
    g["kat_t_starts"] = np.arange(0, 7, dtype=np.float32)
    g["kat_t_ends"] = np.arange(1, 8, dtype=np.float32)
    g["kat_sigmas"] = g["kat_alphas"].copy()
    # printed to 2 significant digits in the docstrings -> compare with atol 6e-3
    g["kat_trans_from_density"] = np.array([1.00, 0.67, 0.30, 1.00, 0.45, 1.00, 1.00], np.float32)
    g["kat_alphas_from_density"] = np.array([0.33, 0.55, 0.095, 0.55, 0.095, 0.00, 0.59], np.float32)
    g["kat_weights_from_density"] = np.array([0.33, 0.37, 0.03, 0.55, 0.04, 0.00, 0.59], np.float32)
    g["kat_visibility"] = np.array([1, 1, 0, 1, 0, 0, 1], np.uint8)  # eps=0.3, alpha_thre=0.2

    np.savez_compressed(OUT, **g)
    print("wrote", OUT, len(g), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
