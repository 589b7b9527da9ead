"""Host-side mirror of the reference operator surface for the hash-grid encoder.

Mirrors examples/radiance_fields/ngp.py:22-315 -- `STE_binary`, `STE_multistep`,
`_grid_encode` / `grid_encode`, `GridEncoder(num_dim, n_features, resolutions_list,
log2_hashmap_size, ste_binary, ste_multistep, add_noise, Q)` with `.params`, `.offsets_list`,
`.resolutions_list`, `.n_output_dims` and the three forwards (`forward`, `forward_diff_levels`,
`forward_given_params`) -- same names, argument meaning and error behaviour, so code written
against the reference keeps working.  Differences are internal and B200-motivated:

  * STE_binary is one fused kernel each way instead of >= 6 elementwise passes over the
    whole table per call (ngp.py:26-31 / SURVEY T1).
  * With `ste_binary=True` the forward gathers from a 1-bit/parameter sign table
    (`cnc_sign_pack`, 1/32 of the fp32 table -> L2 resident); it is re-packed only when
    `params` changed (tensor version counter), not on every forward.  Results are identical
    because every STE value is exactly +-1.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _gridencoder as _backend


class STE_binary(Function):
    """ngp.py:22-39 (== utils_bpp_acc.py:164-181)."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return _backend.ste_binary_forward(input.contiguous())

    @staticmethod
    def backward(ctx, grad_output):
        (input,) = ctx.saved_tensors
        return _backend.ste_binary_backward(input.contiguous(), grad_output.contiguous())


class STE_multistep(Function):
    """ngp.py:41-47."""

    @staticmethod
    def forward(ctx, input, Q):
        return torch.round(input * Q) / Q

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None


def _levels(offsets_list, resolutions_list, min_level_id, n_levels_calc):
    """Slicing rule of ngp.py:86-109: int min level -> sliced lists and no per-point tensor."""
    if isinstance(min_level_id, int):
        max_level_id = min_level_id + n_levels_calc
        return offsets_list[min_level_id:max_level_id + 1], resolutions_list[min_level_id:max_level_id], None
    return offsets_list, resolutions_list, min_level_id


def _prep(inputs, binary_vxl):
    inputs = inputs.contiguous()
    Rb = 128
    if binary_vxl is not None:
        binary_vxl = binary_vxl.contiguous()
        Rb = binary_vxl.shape[-1]
        assert len(binary_vxl.shape) == inputs.shape[-1]
    return inputs, binary_vxl, Rb


class _grid_encode(Function):
    """ngp.py:49-165: fp32-table forward (K1) / scatter-add backward (K2)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets_list, resolutions_list, calc_grad_inputs=False,
                min_level_id=None, n_levels_calc=1, binary_vxl=None, PV=0):
        if calc_grad_inputs:
            print('calc_grad_inputs not applicable!')
            assert False
        inputs, binary_vxl, Rb = _prep(inputs, binary_vxl)
        N, num_dim = inputs.shape
        n_features = embeddings.shape[1]
        embeddings = embeddings.contiguous()
        outputs = torch.empty(n_levels_calc, N, n_features, device=inputs.device, dtype=embeddings.dtype)
        offs, ress, mlid = _levels(offsets_list, resolutions_list, min_level_id, n_levels_calc)
        _backend.grid_encode_forward(inputs, embeddings, offs.contiguous(), ress.contiguous(), outputs, N,
                                     num_dim, n_features, n_levels_calc, 0, Rb, PV, None, binary_vxl, mlid)
        outputs = outputs.permute(1, 0, 2).reshape(N, n_levels_calc * n_features)
        ctx.save_for_backward(inputs, embeddings, offsets_list, resolutions_list, binary_vxl,
                              None if isinstance(min_level_id, int) else min_level_id)
        ctx.dims = [N, num_dim, n_features, n_levels_calc, min_level_id if isinstance(min_level_id, int) else None, Rb]
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets_list, resolutions_list, binary_vxl, mlid_t = ctx.saved_tensors
        N, num_dim, n_features, n_levels_calc, mlid_i, Rb = ctx.dims
        grad = grad.view(N, n_levels_calc, n_features).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        offs, ress, mlid = _levels(offsets_list, resolutions_list, mlid_i if mlid_t is None else mlid_t, n_levels_calc)
        _backend.grid_encode_backward(grad, inputs, embeddings, offs.contiguous(), ress.contiguous(),
                                      grad_embeddings, N, num_dim, n_features, n_levels_calc, 0, Rb, None, None,
                                      binary_vxl, mlid)
        return None, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class _SignCache:
    """1-bit sign table of a parameter tensor.

    The tensor's (address, version, shape) is only trusted as "unchanged" outside training: fused optimizers
    (torch.optim.Adam(fused=True), our own table Adam) write the parameters WITHOUT bumping the autograd version counter,
    so during training the plane is re-packed on every forward (one read pass of the table) -- unless an optimizer that
    keeps the plane up to date itself has `publish`ed it (dp.ShardedTableAdam writes the sign plane in the same pass as the
    update)."""

    def __init__(self):
        self.key = None
        self.bits = None
        self.published = False

    def publish(self, params: torch.Tensor, bits: torch.Tensor) -> None:
        """`bits` is, and will be kept, the sign plane of `params` by whoever updates `params` in place"""
        self.bits, self.key, self.published = bits, (params.data_ptr(), params._version, tuple(params.shape)), True

    def get(self, params: torch.Tensor, trust_version: bool = True) -> torch.Tensor:
        key = (params.data_ptr(), params._version, tuple(params.shape))
        if key != self.key or not (trust_version or self.published):
            self.bits = _backend.sign_pack(params.detach().contiguous(), None)
            self.key, self.published = key, False
        return self.bits


class _grid_encode_ste(Function):
    """STE_binary (ngp.py:244-245) fused with K1/K2: forward reads the sign bit-planes, backward
    scatter-adds and applies the STE mask |p| <= 1 (ngp.py:33-39) in one pass."""

    @staticmethod
    def forward(ctx, inputs, params, bits, offsets_list, resolutions_list, min_level_id, n_levels_calc,
                binary_vxl):
        inputs, binary_vxl, Rb = _prep(inputs, binary_vxl)
        N, num_dim = inputs.shape
        n_features = params.shape[1]
        outputs = torch.empty(n_levels_calc, N, n_features, device=inputs.device, dtype=torch.float32)
        offs, ress, mlid = _levels(offsets_list, resolutions_list, min_level_id, n_levels_calc)
        _backend.grid_encode_forward_bits(inputs, bits, offs.contiguous(), ress.contiguous(), outputs, N, num_dim,
                                          n_features, n_levels_calc, Rb, binary_vxl, mlid)
        outputs = outputs.permute(1, 0, 2).reshape(N, n_levels_calc * n_features)
        ctx.save_for_backward(inputs, params, offsets_list, resolutions_list, binary_vxl,
                              None if isinstance(min_level_id, int) else min_level_id)
        ctx.dims = [N, num_dim, n_features, n_levels_calc, min_level_id if isinstance(min_level_id, int) else None, Rb]
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, params, offsets_list, resolutions_list, binary_vxl, mlid_t = ctx.saved_tensors
        N, num_dim, n_features, n_levels_calc, mlid_i, Rb = ctx.dims
        grad = grad.view(N, n_levels_calc, n_features).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(params)
        offs, ress, mlid = _levels(offsets_list, resolutions_list, mlid_i if mlid_t is None else mlid_t, n_levels_calc)
        _backend.grid_encode_backward(grad, inputs, params, offs.contiguous(), ress.contiguous(), grad_embeddings,
                                      N, num_dim, n_features, n_levels_calc, 0, Rb, None, None, binary_vxl, mlid)
        grad_params = _backend.ste_binary_backward(params.contiguous(), grad_embeddings)
        return None, grad_params, None, None, None, None, None, None


class GridEncoder(nn.Module):
    """ngp.py:171-315."""

    def __init__(self, num_dim=3, n_features=2,
                 resolutions_list=(16, 23, 32, 46, 64, 92, 128, 184, 256, 368, 512, 736),
                 log2_hashmap_size=19, ste_binary=False, ste_multistep=False, add_noise=False, Q=1):
        super().__init__()
        resolutions_list = torch.tensor(np.array(resolutions_list)).to(torch.int)
        n_levels = resolutions_list.numel()
        self.num_dim = num_dim
        self.n_levels = n_levels
        self.n_features = n_features
        self.log2_hashmap_size = log2_hashmap_size
        self.output_dim = n_levels * n_features
        self.ste_binary = ste_binary
        self.ste_multistep = ste_multistep
        self.add_noise = add_noise
        self.Q = Q

        offsets_list, offset = [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(n_levels):
            resolution = resolutions_list[i].item()
            params_in_level = min(self.max_params, resolution ** num_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets_list.append(offset)
            offset += params_in_level
        offsets_list.append(offset)
        self.register_buffer('offsets_list', torch.from_numpy(np.array(offsets_list, dtype=np.int32)))
        self.register_buffer('resolutions_list', resolutions_list)
        self.n_params = offsets_list[-1] * n_features
        self.params = nn.Parameter(torch.empty(offset, n_features))
        self.reset_parameters()
        self.n_output_dims = n_levels * n_features
        self._sign_cache = _SignCache()
        self._sign_cache_out = _SignCache()

    def reset_parameters(self):
        std = 1e-4
        with torch.no_grad():   # (an in-place op on the parameter itself: bumps the version the sign cache is keyed on)
            self.params.uniform_(-std, std)

    def invalidate(self):
        """forget the cached sign planes (after writing to `params` through `.data`, a raw pointer or a fused optimizer)"""
        self._sign_cache.key = self._sign_cache_out.key = None
        self._sign_cache.published = False

    def train(self, mode: bool = True):
        if mode != self.training and not self._sign_cache.published:
            self._sign_cache.key = None      # the last optimizer step may not have bumped the version (see _SignCache)
        return super().train(mode)

    def sign_bits(self) -> torch.Tensor:
        """the 1-bit plane of `params` the forward gathers from"""
        return self._sign_cache.get(self.params, trust_version=not self.training)

    def __repr__(self):
        return (f"GridEncoder: num_dim={self.num_dim} n_levels={self.n_levels} n_features={self.n_features} "
                f"resolutions={self.resolutions_list.tolist()} params={tuple(self.params.shape)} "
                f"ste_binary={self.ste_binary}")

    # -- embedding selection, ngp.py:239-252 --------------------------------------------------
    def _encode(self, inputs, outspace_params, min_level_id, n_levels_calc, test_phase, binary_vxl, PV):
        params = self.params if outspace_params is None else outspace_params
        if self.ste_binary:
            if outspace_params is None:
                bits = self.sign_bits()
            else:
                # a caller-provided table (decode: the partially reconstructed one; a fresh STE output per call): address and
                # version do not identify it -- the allocator hands the same block to the next temporary -- so no caching
                bits = _backend.sign_pack(params.detach().contiguous(), None)
            return _grid_encode_ste.apply(inputs, params, bits, self.offsets_list, self.resolutions_list,
                                          min_level_id, n_levels_calc, binary_vxl)
        if self.add_noise and not test_phase:
            embeddings = params + (torch.rand_like(params) - 0.5) * (1 / self.Q)
        elif self.ste_multistep or (self.add_noise and test_phase):
            embeddings = STE_multistep.apply(params, self.Q)
        else:
            embeddings = params
        return grid_encode(inputs, embeddings, self.offsets_list, self.resolutions_list, False, min_level_id,
                           n_levels_calc, binary_vxl, PV)

    def forward(self, inputs, min_level_id=None, max_level_id=None, test_phase=False, outspace_params=None,
                binary_vxl=None, PV=0):
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.num_dim)
        min_level_id = 0 if min_level_id is None else max(min_level_id, 0)
        max_level_id = self.n_levels if max_level_id is None else min(max_level_id, self.n_levels)
        n_levels_calc = max_level_id - min_level_id
        outputs = self._encode(inputs, outspace_params, min_level_id, n_levels_calc, test_phase, binary_vxl, PV)
        return outputs.view(prefix_shape + [n_levels_calc * self.n_features])

    def forward_diff_levels(self, inputs, min_level_id_list=None, n_levels_calc=1, test_phase=False,
                            outspace_params=None, binary_vxl=None, PV=0):
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.num_dim)
        outputs = self._encode(inputs, outspace_params, min_level_id_list.contiguous(), n_levels_calc, test_phase,
                               binary_vxl, PV)
        return outputs.view(prefix_shape + [n_levels_calc * self.n_features])

    def forward_given_params(self, inputs, offsets_list, resolutions_list, outspace_params=None, binary_vxl=None,
                             PV=0):
        assert inputs.shape[-1] == 2
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, 2)
        outputs = grid_encode(inputs, outspace_params, offsets_list, resolutions_list, False, 0, 1, binary_vxl, PV)
        return outputs.view(prefix_shape + [self.n_features])
