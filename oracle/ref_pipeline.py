"""Reference-faithful GPU pipeline around the UNMODIFIED reference kernels in oracle/_ref.

TEST INFRASTRUCTURE ONLY (used by tests/ and by `bench.py --impl reference`).  /root/reference is
not available on the GPU box, so the reference's *Python* data flow is restated here op for op
(the kernels themselves are the reference's own binaries):

  STE_binary as 6 eager elementwise passes         examples/radiance_fields/ngp.py:22-31
  _grid_encode.forward: empty [L,N,F] -> K1 -> permute(1,0,2).reshape     ngp.py:52-116
  _grid_encode.backward: permute+contiguous, zeros_like, K2               ngp.py:121-165
  compose_3D_2D_embed: 4 encoders + torch.cat + Embedder (21 small ops)   ngp.py:569-645
  NGPRadianceField_mygrid_2D3D.query_density / _query_rgb / forward       ngp.py:514-566
tinycudann's SphericalHarmonics (third party, absent) is replaced by the same polynomial in
torch ops with an fp16 round trip.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ref_ext


class _STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        input = torch.clamp(input, min=-1, max=1)
        p = (input >= 0) * (+1.0)
        n = (input < 0) * (-1.0)
        return p + n

    @staticmethod
    def backward(ctx, grad_output):
        (input,) = ctx.saved_tensors
        i2 = input.clone().detach()
        i3 = torch.clamp(i2, -1, 1)
        mask = (i3 == i2) + 0.0
        return grad_output * mask


class _RefGridEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets_list, resolutions_list, min_level_id, n_levels_calc, binary_vxl):
        be = ref_ext.load("_gridencoder")
        inputs = inputs.contiguous()
        Rb = 128 if binary_vxl is None else binary_vxl.shape[-1]
        N, D = inputs.shape
        F = embeddings.shape[1]
        outputs = torch.empty(n_levels_calc, N, F, device=inputs.device, dtype=embeddings.dtype)
        mx = min_level_id + n_levels_calc
        be.grid_encode_forward(inputs, embeddings, offsets_list[min_level_id:mx + 1],
                               resolutions_list[min_level_id:mx], outputs, N, D, F, n_levels_calc, 0, Rb, 0.0,
                               None, binary_vxl, None)
        outputs = outputs.permute(1, 0, 2).reshape(N, n_levels_calc * F)
        ctx.save_for_backward(inputs, embeddings, offsets_list, resolutions_list, binary_vxl)
        ctx.dims = [N, D, F, n_levels_calc, min_level_id, mx, Rb]
        return outputs

    @staticmethod
    def backward(ctx, grad):
        be = ref_ext.load("_gridencoder")
        inputs, embeddings, offsets_list, resolutions_list, binary_vxl = ctx.saved_tensors
        N, D, F, L, mn, mx, Rb = ctx.dims
        grad = grad.view(N, L, F).permute(1, 0, 2).contiguous()
        ge = torch.zeros_like(embeddings)
        be.grid_encode_backward(grad, inputs, embeddings, offsets_list[mn:mx + 1], resolutions_list[mn:mx], ge, N, D,
                                F, L, 0, Rb, None, None, binary_vxl, None)
        return None, ge, None, None, None, None, None


class RefGridEncoder(nn.Module):
    def __init__(self, num_dim, n_features, resolutions_list, log2_hashmap_size):
        super().__init__()
        res = torch.tensor(np.array(resolutions_list)).to(torch.int)
        offs, off = [], 0
        for r in res.tolist():
            n = int(np.ceil(min(2 ** log2_hashmap_size, r ** num_dim) / 8) * 8)
            offs.append(off)
            off += n
        offs.append(off)
        self.num_dim, self.n_features, self.n_levels = num_dim, n_features, len(resolutions_list)
        self.register_buffer("offsets_list", torch.from_numpy(np.array(offs, dtype=np.int32)))
        self.register_buffer("resolutions_list", res)
        self.params = nn.Parameter(torch.empty(off, n_features).uniform_(-1e-4, 1e-4))
        self.n_output_dims = self.n_levels * n_features

    def forward(self, inputs, min_level_id=None, max_level_id=None, binary_vxl=None):
        emb = _STE.apply(self.params)
        mn = 0 if min_level_id is None else max(min_level_id, 0)
        mx = self.n_levels if max_level_id is None else min(max_level_id, self.n_levels)
        return _RefGridEncode.apply(inputs.view(-1, self.num_dim), emb, self.offsets_list, self.resolutions_list, mn,
                                    mx - mn, binary_vxl)


def _sh16_torch(d01):
    x, y, z = (d01 * 2 - 1).unbind(-1)
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z,
         -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
         0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
         0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
         2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
         0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
         1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, -1).half()  # tcnn returns fp16; torch.cat promotes it back (ngp.py:541-542)


def _embed(x):
    outs = [x]
    for f in (2.0 ** torch.linspace(0.0, 9.0, steps=10)).tolist():
        outs += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(outs, -1)


class RefField(nn.Module):
    """NGPRadianceField_mygrid_2D3D restated around the reference kernels (state_dict keys match
    cnc_b200.field.NGPRadianceField_mygrid_2D3D so weights can be copied across)."""

    def __init__(self, aabb, resolutions_list, log2_hashmap_size, resolutions_list_2D, log2_hashmap_size_2D,
                 n_features_per_level=8, n_neurons=160):
        super().__init__()
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=torch.float32))
        F = n_features_per_level
        self.geo_feat_dim = min(127, max(15, F * 10 - 1))
        self.encoding_xyz = RefGridEncoder(3, F, resolutions_list, log2_hashmap_size)
        self.encoding_xy = RefGridEncoder(2, F, resolutions_list_2D, log2_hashmap_size_2D)
        self.encoding_xz = RefGridEncoder(2, F, resolutions_list_2D, log2_hashmap_size_2D)
        self.encoding_yz = RefGridEncoder(2, F, resolutions_list_2D, log2_hashmap_size_2D)
        in_ch = self.encoding_xyz.n_output_dims + 3 * self.encoding_xy.n_output_dims + 63
        self.network = nn.Sequential(nn.Linear(in_ch, n_neurons), nn.ReLU(inplace=True),
                                     nn.Linear(n_neurons, 1 + self.geo_feat_dim))
        self.mlp_head = nn.Sequential(nn.Linear(16 + self.geo_feat_dim, n_neurons), nn.ReLU(inplace=True),
                                      nn.Linear(n_neurons, n_neurons), nn.ReLU(inplace=True),
                                      nn.Linear(n_neurons, 3))

    def load_from(self, field):
        """copy weights from a cnc_b200.field.NGPRadianceField_mygrid_2D3D"""
        with torch.no_grad():
            for k in ("xyz", "xy", "xz", "yz"):
                getattr(self, f"encoding_{k}").params.copy_(getattr(field.mlp_base, f"encoding_{k}").params)
            self.network.load_state_dict(field.mlp_base.network.state_dict())
            self.mlp_head.load_state_dict(field.mlp_head.state_dict())

    def mlp_base(self, x):
        x_x, y_y, z_z = torch.chunk(x, 3, dim=-1)
        out_xyz = self.encoding_xyz(x)
        out_xy = self.encoding_xy(torch.cat([x_x, y_y], dim=-1))
        out_xz = self.encoding_xz(torch.cat([x_x, z_z], dim=-1))
        out_yz = self.encoding_yz(torch.cat([y_y, z_z], dim=-1))
        out_i = torch.cat([out_xyz, out_xy, out_xz, out_yz], dim=-1)
        out_i = torch.cat([out_i, _embed(x)], dim=-1)
        return self.network(out_i)

    def query_density(self, x, return_feat=False):
        aabb_min, aabb_max = torch.split(self.aabb, 3, dim=-1)
        x = (x - aabb_min) / (aabb_max - aabb_min)
        selector = ((x > 0.0) & (x < 1.0)).all(dim=-1)
        x = self.mlp_base(x.view(-1, 3)).view(list(x.shape[:-1]) + [1 + self.geo_feat_dim]).to(x)
        d, feat = torch.split(x, [1, self.geo_feat_dim], dim=-1)
        density = torch.exp(d - 1) * selector[..., None]
        return (density, feat) if return_feat else density

    def forward(self, positions, directions):
        density, embedding = self.query_density(positions, return_feat=True)
        d = _sh16_torch(((directions + 1.0) / 2.0).reshape(-1, 3))
        h = torch.cat([d, embedding.reshape(-1, self.geo_feat_dim)], dim=-1)
        rgb = torch.sigmoid(self.mlp_head(h).reshape(list(embedding.shape[:-1]) + [3]))
        return rgb, density
